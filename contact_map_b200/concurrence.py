"""Contact concurrences on the GPU.

Mirror of /root/reference/contact_map/concurrence.py (Concurrence :16-51,
_regularize_contact_input :66-97, AtomContactConcurrence :100-125, ResidueContactConcurrence
:128-162).  The reference calls ``md.compute_distances`` once per contact and thresholds the
float32 distances in Python; here one kernel (``cmb_group_contacts``) evaluates every contact of
every frame: one CTA per frame, the atoms the contacts touch staged in shared memory.  Plotting
(``ConcurrencePlotter``) is out of scope.
"""
import itertools

import numpy as np

from .contact_count import ContactCount
from .contact_map import ContactObject, _split_trajectory
from .min_dist import _is_orthogonal, _to_device


class _BoolRows(object):
    """``[n_contacts][n_frames]`` booleans behaving like the reference's list of lists (rows are
    materialised as Python lists on access; the array itself is ``.array``)."""

    def __init__(self, array):
        self.array = np.asarray(array, dtype=bool)

    def __len__(self):
        return self.array.shape[0]

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            return [row.tolist() for row in self.array[idx]]
        return self.array[idx].tolist()

    def __iter__(self):
        return (row.tolist() for row in self.array)

    def __eq__(self, other):
        return list(self) == list(other)

    def __ne__(self, other):
        return not self == other


class Concurrence(object):
    """Superclass for contact concurrence objects (reference :16-51): ``values[contact][frame]``
    tells whether the contact is present, ``labels`` names the contacts."""

    def __init__(self, values, labels=None):
        self.values = values
        self.labels = labels

    def set_labels(self, labels):
        self.labels = labels

    def __getitem__(self, label):
        idx = self.labels.index(label)
        return self.values[idx]


def _regularize_contact_input(contact_input, atom_or_res):
    """``ContactObject`` / ``ContactCount`` / ``most_common()`` list -> ``most_common()`` list
    (reference :66-97)."""
    if isinstance(contact_input, ContactObject):
        contact_input = contact_input.contacts[atom_or_res]
    if isinstance(contact_input, ContactCount):
        contact_input = contact_input.most_common()
    return contact_input


def _group_contacts(trajectory, pair_lists, cutoff):
    from .engine import get_engine
    engine = get_engine()
    xyz, box = _split_trajectory(trajectory)
    d_xyz, d_box = _to_device(xyz, box, engine)
    sizes = [len(p) for p in pair_lists]
    ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    pairs = (np.concatenate([np.asarray(p, dtype=np.int64).reshape(-1, 2) for p in pair_lists])
             if sum(sizes) else np.zeros((0, 2), dtype=np.int64))
    return engine.group_contacts(d_xyz, d_box, pairs, ptr, cutoff, _is_orthogonal(box))


class AtomContactConcurrence(Concurrence):
    """Contact concurrences for atom contacts (reference :100-125): ``values[c][f]`` is
    ``compute_distances(pair c)[f] < cutoff``."""

    def __init__(self, trajectory, atom_contacts, cutoff=0.45):
        atom_contacts = _regularize_contact_input(atom_contacts, "atom")
        atom_pairs = [[[contact[0][0].index, contact[0][1].index]] for contact in atom_contacts]
        labels = [str(contact[0]) for contact in atom_contacts]
        values = _BoolRows(_group_contacts(trajectory, atom_pairs, cutoff))
        super(AtomContactConcurrence, self).__init__(values=values, labels=labels)


class ResidueContactConcurrence(Concurrence):
    """Contact concurrences for residue contacts (reference :128-162): the minimum distance over
    the product of the two residues' selected atoms, compared with the cutoff."""

    def __init__(self, trajectory, residue_contacts, cutoff=0.45, select="and symbol != 'H'"):
        residue_contacts = _regularize_contact_input(residue_contacts, "residue")
        labels = [str(contact[0]) for contact in residue_contacts]
        topology = trajectory.topology
        cache = {}

        def select_residue(idx):
            if idx not in cache:
                cache[idx] = _select_residue_atoms(topology, idx, select)
            return cache[idx]

        pair_lists = [list(itertools.product(select_residue(contact[0][0].index),
                                             select_residue(contact[0][1].index)))
                      for contact in residue_contacts]
        values = _BoolRows(_group_contacts(trajectory, pair_lists, cutoff))
        super(ResidueContactConcurrence, self).__init__(values=values, labels=labels)


def _select_residue_atoms(topology, res_index, select):
    """``topology.select("resid <i> " + select)`` (reference :147-149).  MDTraj topologies take the
    full expression; the built-in Topology evaluates the two forms the reference uses."""
    expression = "resid " + str(res_index) + " " + select
    try:
        return [int(a) for a in topology.select(expression)]
    except ValueError:
        atoms = [int(a) for a in topology.select("resid " + str(res_index))]
        extra = " ".join(select.split())
        if extra == "":
            return atoms
        if extra == "and symbol != 'H'":
            heavy = set(int(a) for a in topology.select("symbol != 'H'"))
            return [a for a in atoms if a in heavy]
        raise
