"""GPU parity: contact concurrences against vectors produced by the unmodified
reference (tests/golden/make_golden_concurrence.py) and the literal expectations of the
reference's own test-suite (contact_map/tests/test_concurrence.py:88-95,125-134)."""
import os
import warnings

import numpy as np
import pytest

import contact_map_b200 as cm
from contact_map_b200.concurrence import _regularize_contact_input
from contact_map_b200.trajectory import Topology, Trajectory

pytestmark = pytest.mark.gpu
warnings.simplefilter("ignore", PendingDeprecationWarning)
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["conc_pdb", "conc_syn_ortho", "conc_syn_tric", "conc_syn_open"]


def _load(name):
    case = dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
    top = Topology(case["atom_residue"], case["residue_chain"],
                   atom_names=[str(s) for s in case["atom_names"]],
                   atom_elements=[str(s) for s in case["atom_elements"]],
                   residue_names=[str(s) for s in case["residue_names"]],
                   residue_resseq=case["residue_resseq"])
    return case, Trajectory(case["xyz"], top, case.get("box"))


@pytest.mark.parametrize("name", CASES)
def test_concurrences_match_reference(engine, name):
    case, traj = _load(name)
    top = traj.topology
    cutoff = float(case["cutoff"])
    atom_contacts = [([top.atom(int(i)), top.atom(int(j))], 0.0) for i, j in case["atom_contact_pairs"]]
    res_contacts = [([top.residue(int(i)), top.residue(int(j))], 0.0) for i, j in case["res_contact_pairs"]]
    atoms = cm.AtomContactConcurrence(traj, atom_contacts, cutoff=cutoff)
    assert atoms.labels == [str(s) for s in case["atom_labels"]]
    assert np.array_equal(atoms.values.array, case["atom_values"])
    heavy = cm.ResidueContactConcurrence(traj, res_contacts, cutoff=cutoff)
    every = cm.ResidueContactConcurrence(traj, res_contacts, cutoff=cutoff, select="")
    assert heavy.labels == [str(s) for s in case["res_labels"]]
    assert np.array_equal(heavy.values.array, case["res_values_heavy"])
    assert np.array_equal(every.values.array, case["res_values_all"])
    # rows behave like the reference's lists
    assert atoms[atoms.labels[0]] == case["atom_values"][0].tolist()
    assert len(atoms.values) == len(atom_contacts) and atoms.values == case["atom_values"].tolist()
    # device-resident coordinates give the same answer
    import torch

    class DeviceTraj(object):
        topology = top
        xyz = torch.as_tensor(traj.xyz, device=engine.device)
        unitcell_vectors = (None if traj.unitcell_vectors is None
                            else torch.as_tensor(traj.unitcell_vectors, device=engine.device))

        def __len__(self):
            return self.xyz.shape[0]

    dev = cm.AtomContactConcurrence(DeviceTraj(), atom_contacts, cutoff=cutoff)
    assert np.array_equal(dev.values.array, case["atom_values"])


def test_reference_suite_expectations(engine):
    """test_concurrence.py: contacts from ContactFrequency(...).most_common(), literal tables."""
    case, traj = _load("conc_pdb")
    top = traj.topology
    query = top.select("resSeq 3")
    haystack = top.select("resSeq 1 to 2")
    contacts = cm.ContactFrequency(traj, query, haystack, cutoff=0.051, n_neighbors_ignored=0)
    most_common = contacts.atom_contacts.most_common()
    assert _regularize_contact_input(most_common, "atom") == most_common
    assert _regularize_contact_input(contacts.atom_contacts, "atom") == most_common
    assert _regularize_contact_input(contacts, "atom") == most_common
    with pytest.raises(RuntimeError):
        _regularize_contact_input(contacts, "foo")

    label_to_pair = {'[AAA1-H, LLL3-H]': 'AH-LH', '[LLL3-H, AAA1-H]': 'AH-LH',
                     '[AAA1-C1, LLL3-C1]': 'AC1-LC1', '[LLL3-C1, AAA1-C1]': 'AC1-LC1',
                     '[BBB2-H, LLL3-C2]': 'BH-LC2', '[LLL3-C2, BBB2-H]': 'BH-LC2',
                     '[AAA1-C2, LLL3-C2]': 'AC2-LC2', '[LLL3-C2, AAA1-C2]': 'AC2-LC2',
                     '[AAA1-C2, LLL3-C1]': 'AC2-LC1', '[LLL3-C1, AAA1-C2]': 'AC2-LC1',
                     '[BBB2-C2, LLL3-C2]': 'BC2-LC2', '[LLL3-C2, BBB2-C2]': 'BC2-LC2'}
    expected = {'AH-LH': [False, False, True, True, False], 'AC1-LC1': [False, True, False, False, False],
                'BH-LC2': [False, False, False, True, False], 'AC2-LC2': [False, True, False, False, False],
                'AC2-LC1': [True, False, False, False, False], 'BC2-LC2': [True, False, False, False, False]}
    conc = cm.AtomContactConcurrence(trajectory=traj, atom_contacts=most_common, cutoff=0.051)
    assert len(conc.labels) == len(label_to_pair) / 2
    for label in conc.labels:
        assert conc[label] == expected[label_to_pair[label]]
    conc.set_labels([label_to_pair[label] for label in conc.labels])
    for label in conc.labels:
        assert conc[label] == expected[label]

    res_label_to_pair = {'[AAA1, LLL3]': 'AL', '[LLL3, AAA1]': 'AL', '[BBB2, LLL3]': 'BL', '[LLL3, BBB2]': 'BL'}
    res_expected = {'heavy': {'AL': [True, True, False, False, False], 'BL': [True, False, False, False, False]},
                    'all': {'AL': [True, True, True, True, False], 'BL': [True, False, False, True, False]}}
    res_common = contacts.residue_contacts.most_common()
    for kind, select in (("heavy", "and symbol != 'H'"), ("all", "")):
        rc = cm.ResidueContactConcurrence(trajectory=traj, residue_contacts=res_common, cutoff=0.051,
                                          select=select)
        assert len(rc.labels) == 2
        for label in rc.labels:
            assert rc[label] == res_expected[kind][res_label_to_pair[label]]
