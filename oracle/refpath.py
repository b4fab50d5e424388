"""Plain-Python restatement of the reference's CPU path (ORACLE -- test infrastructure only).

This is the algorithm of /root/reference/contact_map as the reference executes it on the
CPU: a Python loop over frames around a native neighbour list, followed by per-residue set
algebra and ``collections.Counter`` accumulation.  It is (a) a second, independent check of
the closed-form predicate in ``cm_oracle.c`` and (b) the CPU baseline timed by ``bench.py``
(kind "port": the reference's MDTraj neighbour list is replaced by ``oracle/cm_oracle.c``).

Reference lines followed:
  ContactObject.__init__              contact_map/contact_map.py:120-150
  residue_neighborhood                contact_map/contact_map.py:41-60
  _residue_ignore_atom_idxs           contact_map/contact_map.py:442-461
  AtomSlicedIndexer / _atom_slice     contact_map/atom_indexer.py:5-78
  ContactObject._contact_map          contact_map/contact_map.py:556-615
  ContactFrequency._build_contact_map contact_map/contact_map.py:717-744
  _build_contacts                     contact_map/contact_trajectory.py:7-44
  MinimumDistanceCounter              contact_map/min_dist.py:134-168
  NearestAtoms._calculate_nearest     contact_map/min_dist.py:47-70
  default_slices / reduce_all_results contact_map/frequency_task.py:27-67,125-141
"""
import collections

import numpy as np

from . import clib


class SystemTables(object):
    """Index bookkeeping of one (topology, query, haystack, n_ignored) combination."""

    def __init__(self, atom_res, res_chain, query, haystack, n_ignored):
        atom_res = np.asarray(atom_res, dtype=np.int64)
        res_chain = np.asarray(res_chain, dtype=np.int64)
        self.n_atoms = atom_res.shape[0]
        self.query = set(int(q) for q in query)
        self.haystack = set(int(h) for h in haystack)
        # contact_map.py:138-139 -- sorted union; atom_indexer.py:37-44 -- rank = sliced index
        self.all_atoms = sorted(self.query | self.haystack)
        self.sliced_of = {real: pos for pos, real in enumerate(self.all_atoms)}
        self.res_of_sliced = [int(atom_res[real]) for real in self.all_atoms]
        self.sliced_query = set(self.sliced_of[q] for q in self.query)
        self.sliced_haystack = set(self.sliced_of[h] for h in self.haystack)
        # atom_indexer.py:24-29 -- query atoms grouped by residue
        self.res_query_atoms = collections.defaultdict(list)
        for s in self.sliced_query:
            self.res_query_atoms[self.res_of_sliced[s]].append(s)
        # contact_map.py:41-60,442-461 -- +-n window clipped to the residue's own chain
        atoms_of_res = collections.defaultdict(list)
        for real, r in enumerate(atom_res):
            atoms_of_res[int(r)].append(real)
        n_res = res_chain.shape[0]
        self.ignore = {}
        for r in self.res_query_atoms:
            window = [k for k in range(r - n_ignored, r + n_ignored + 1)
                      if 0 <= k < n_res and res_chain[k] == res_chain[r]]
            ignored = set()
            for k in window:
                ignored.update(self.sliced_of[a] for a in atoms_of_res[k] if a in self.sliced_of)
            self.ignore[r] = ignored
        self.atom_res = atom_res
        self.res_chain = res_chain


def frame_contacts(tables, xyz_used_frame, box_frame, cutoff):
    """One frame: (set of sliced atom pairs, set of residue pairs) -- contact_map.py:575-615."""
    neighbours = clib.neighborlist(xyz_used_frame, box_frame, cutoff)
    atom_pairs, residue_pairs = set(), set()
    for res, q_atoms in tables.res_query_atoms.items():
        ignored = tables.ignore[res]
        for a in q_atoms:
            partners = (set(int(j) for j in neighbours[a]) - ignored) & tables.sliced_haystack
            atom_pairs.update(frozenset((a, b)) for b in partners)
            residue_pairs.update(frozenset((res, tables.res_of_sliced[b])) for b in partners)
    return atom_pairs, residue_pairs


def contact_frequency(xyz, box, atom_res, res_chain, query, haystack, cutoff, n_ignored):
    """(atom Counter, residue Counter) of integer frame counts in REAL indices."""
    tables = SystemTables(atom_res, res_chain, query, haystack, n_ignored)
    used = np.asarray(tables.all_atoms, dtype=np.int64)
    xyz_used = np.ascontiguousarray(np.asarray(xyz, dtype=np.float32)[:, used])
    atom_counts, residue_counts = collections.Counter(), collections.Counter()
    for f in range(xyz_used.shape[0]):
        atoms, residues = frame_contacts(tables, xyz_used[f], None if box is None else box[f], cutoff)
        atom_counts.update(atoms)
        residue_counts.update(residues)
    real = tables.all_atoms
    atom_counts = collections.Counter({frozenset(real[s] for s in pair): v
                                       for pair, v in atom_counts.items()})
    return atom_counts, residue_counts


def contact_trajectory(xyz, box, atom_res, res_chain, query, haystack, cutoff, n_ignored):
    """Per-frame (atom pair set, residue pair set) in REAL indices."""
    tables = SystemTables(atom_res, res_chain, query, haystack, n_ignored)
    used = np.asarray(tables.all_atoms, dtype=np.int64)
    xyz_used = np.ascontiguousarray(np.asarray(xyz, dtype=np.float32)[:, used])
    real = tables.all_atoms
    out = []
    for f in range(xyz_used.shape[0]):
        atoms, residues = frame_contacts(tables, xyz_used[f], None if box is None else box[f], cutoff)
        out.append((set(frozenset(real[s] for s in pair) for pair in atoms), residues))
    return out


def role_arrays(n_atoms, query, haystack):
    role = np.zeros(n_atoms, dtype=np.uint8)
    role[np.asarray(sorted(query), dtype=np.int64)] |= 1
    role[np.asarray(sorted(haystack), dtype=np.int64)] |= 2
    return role


def closed_form_frequency(xyz, box, atom_res, res_chain, query, haystack, cutoff, n_ignored,
                          cells=False):
    """Same result as :func:`contact_frequency` through the closed-form C enumeration."""
    atom_res = np.asarray(atom_res, dtype=np.int64)
    res_chain = np.asarray(res_chain, dtype=np.int64)
    used = np.asarray(sorted(set(int(q) for q in query) | set(int(h) for h in haystack)), dtype=np.int64)
    role = role_arrays(atom_res.shape[0], query, haystack)[used]
    res_u, chain_u = atom_res[used], res_chain[atom_res[used]]
    xyz_used = np.ascontiguousarray(np.asarray(xyz, dtype=np.float32)[:, used])
    atom_counts, residue_counts = collections.Counter(), collections.Counter()
    for f in range(xyz_used.shape[0]):
        pairs = clib.contacts_frame(xyz_used[f], None if box is None else box[f], cutoff, res_u,
                                    chain_u, role, n_ignored, cells=cells)
        atom_counts.update(frozenset((int(used[a]), int(used[b]))) for a, b in pairs)
        residue_counts.update(set(frozenset((int(res_u[a]), int(res_u[b]))) for a, b in pairs))
    return atom_counts, residue_counts


def minimum_distances(xyz, box, query, haystack, orthogonal=True):
    """(min distance float32 [F], first arg-min pair index [F]) -- min_dist.py:138,165-167,
    through the materialised [F, P] array exactly like the reference."""
    pairs = np.array([(q, h) for q in query for h in haystack], dtype=np.int32)
    d = clib.distances(xyz, pairs, box, orthogonal)
    return d.min(axis=1), d.argmin(axis=1)


def nearest_atoms(xyz_frame, box_frame, cutoff, excluded, orthogonal=True):
    """min_dist.py:54-69 for one frame; ``excluded`` maps atom -> iterable of excluded atoms."""
    neighbours = clib.neighborlist(xyz_frame, box_frame, cutoff)
    nearest, nearest_distance = {}, {}
    x = np.asarray(xyz_frame, dtype=np.float32)[np.newaxis]
    b = None if box_frame is None else np.asarray(box_frame, dtype=np.float32)[np.newaxis]
    for atom, neigh in enumerate(neighbours):
        allowed = [int(n) for n in neigh if int(n) not in excluded[atom]]
        if not allowed:
            continue
        pairs = np.array([(atom, n) for n in allowed], dtype=np.int32)
        dist = clib.distances(x, pairs, b, orthogonal)[0]
        best = sorted(zip(dist.tolist(), allowed))[0]
        nearest[atom] = best[1]
        nearest_distance[atom] = np.float32(best[0])
    return nearest, nearest_distance


def block_slices(n_total, n_per_block):
    """frequency_task.py:27-47."""
    full = n_total // n_per_block
    out = [slice(k * n_per_block, (k + 1) * n_per_block) for k in range(full)]
    if n_total % n_per_block:
        out.append(slice(full * n_per_block, n_total))
    return out


def default_slices(n_total, n_workers):
    """frequency_task.py:49-67."""
    return block_slices(n_total, max(1, n_total // n_workers))


def _chunk_job(args):
    xyz, box, atom_res, res_chain, query, haystack, cutoff, n_ignored = args
    return contact_frequency(xyz, box, atom_res, res_chain, query, haystack, cutoff, n_ignored)


def chunked_contact_frequency(xyz, box, atom_res, res_chain, query, haystack, cutoff, n_ignored,
                              n_workers, pool=None):
    """Dask-equivalent runner: default_slices -> map (one ContactFrequency per block) -> left
    fold of Counter additions (frequency_task.py:125-141).  ``pool`` is any object with a
    ``map`` method (e.g. ``multiprocessing.Pool``); serial when None."""
    jobs = [(xyz[s], None if box is None else box[s], atom_res, res_chain, query, haystack,
             cutoff, n_ignored) for s in default_slices(xyz.shape[0], n_workers)]
    results = list(pool.map(_chunk_job, jobs)) if pool is not None else [_chunk_job(j) for j in jobs]
    atoms, residues = results[0]
    for a, r in results[1:]:
        atoms += a
        residues += r
    return atoms, residues
