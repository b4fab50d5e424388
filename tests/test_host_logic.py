"""CPU: host-side logic that needs no GPU (partitioning, containers, reduce plumbing)."""
import collections
import ctypes
import os
import re
import warnings

import numpy as np
import pytest

import contact_map_b200 as cm
from contact_map_b200 import frequency_task, parallel, synth
from contact_map_b200.contact_count import ContactCount, PairTable
from tests import goldens

warnings.simplefilter("ignore", PendingDeprecationWarning)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_block_and_default_slices():
    """Known-answer tables of the reference (test_frequency_task.py:13-33)."""
    assert frequency_task.block_slices(100, 25) == [slice(0, 25), slice(25, 50), slice(50, 75), slice(75, 100)]
    assert frequency_task.block_slices(75, 25) == [slice(0, 25), slice(25, 50), slice(50, 75)]
    assert frequency_task.block_slices(76, 25) == [slice(0, 25), slice(25, 50), slice(50, 75), slice(75, 76)]
    assert frequency_task.default_slices(100, 3) == [slice(0, 33), slice(33, 66), slice(66, 99), slice(99, 100)]
    assert frequency_task.default_slices(77, 3) == [slice(0, 25), slice(25, 50), slice(50, 75), slice(75, 77)]
    assert frequency_task.default_slices(2, 20) == [slice(0, 1), slice(1, 2)]
    from oracle import refpath
    for n, w in ((100, 3), (77, 3), (2, 20), (10000, 8), (5, 5), (1, 4)):
        assert frequency_task.default_slices(n, w) == refpath.default_slices(n, w)


def test_rank_slices_cover_all_frames():
    for n_frames, world in ((100, 2), (77, 3), (10000, 8), (5, 8)):
        seen = []
        for rank in range(world):
            for s in parallel.rank_slices(n_frames, None, rank, world):
                seen.extend(range(s.start, s.stop))
        assert sorted(seen) == list(range(n_frames))


def test_windowed_iterator_table():
    """Known answers of the reference (test_contact_trajectory.py:260-285)."""
    it = cm.WindowedIterator(length=10, width=3, step=2, slow_build=False)
    got = [(a.start, a.stop, s.start, s.stop) for a, s in it]
    assert got == [(0, 3, 0, 0), (3, 5, 0, 2), (5, 7, 2, 4), (7, 9, 4, 6)]
    it = cm.WindowedIterator(length=10, width=3, step=2, slow_build=True)
    got = [(a.start, a.stop, s.start, s.stop) for a, s in it]
    assert got[0] == (0, 2, 0, 0) and got[1] == (2, 4, 0, 1)


def test_pair_table_combine_matches_counter_arithmetic():
    rng = np.random.default_rng(3)
    def rand_counter(k):
        c = collections.Counter()
        for _ in range(k):
            a, b = sorted(rng.choice(30, size=2, replace=False).tolist())
            c[frozenset((a, b))] += int(rng.integers(1, 5))
        return c
    for _ in range(20):
        x, y = rand_counter(40), rand_counter(40)
        tx, ty = PairTable.from_counter(x), PairTable.from_counter(y)
        assert tx.combine(ty, +1).to_counter() == x + y
        assert tx.combine(ty, -1).to_counter() == x - y
        assert tx.scaled(7).to_counter() == collections.Counter({k: float(v) / 7 for k, v in x.items()})


def test_contact_count_container():
    counter = collections.Counter({frozenset([0, 8]): 0.2, frozenset([1, 4]): 0.8, frozenset([4, 6]): 1.0})
    cc = ContactCount(counter, lambda i: i, 10, 10)
    assert cc.n_x == 10 and cc.n_x.min == 0 and cc.n_x.max == 10 and cc.max_size == 10
    assert cc.total_range == (0, 9)
    assert cc.most_common_idx()[0] == (frozenset([4, 6]), 1.0)
    assert cc.filter([1, 4, 6]).counter == collections.Counter({frozenset([1, 4]): 0.8, frozenset([4, 6]): 1.0})
    dense = cc.sparse_matrix.toarray()
    assert dense[0, 8] == 0.2 and dense[8, 0] == 0.2 and dense.shape == (10, 10)
    assert cc.df.shape == (10, 10)
    auto = ContactCount(counter, lambda i: i)
    # low keys span 0..4 (5), high keys span 4..8 (5): not larger -> x = low, y = high
    assert auto.n_x == (0, 5) and auto.n_y == (4, 9)
    with pytest.raises(ValueError):
        ContactCount(counter, lambda i: i, n_x=3)


def test_objects_without_gpu():
    """from_contacts / ContactDifference / add work on finished counters without any device."""
    case = goldens.load("pdb_c0075_n0")
    traj = goldens.trajectory(case)
    atoms = goldens.counter(case["atom_pairs"], case["atom_counts"])
    residues = goldens.counter(case["res_pairs"], case["res_counts"])
    m = cm.ContactFrequency.from_contacts(atoms, residues, n_frames=5, topology=traj.topology,
                                          cutoff=0.075, n_neighbors_ignored=0)
    assert m.atom_contacts.counter[frozenset([1, 4])] == 0.8
    assert m.query_range == (0, 10) and m.haystack_residue_range == (0, 5)
    assert m.use_atom_slice is False
    assert m.most_common_atoms_for_residue(2)[0][1] == 1.0
    assert len(m.most_common_atoms_for_contact([2, 3])) == 4
    ignore = m._residue_ignore_atom_idxs
    assert ignore[2] == {4, 5}
    m2 = cm.ContactFrequency.from_contacts(atoms, residues, n_frames=5, topology=traj.topology,
                                           cutoff=0.075, n_neighbors_ignored=0)
    m.add_contact_frequency(m2)
    assert m.n_frames == 10 and m._atom_contacts[frozenset([1, 4])] == 8
    diff = m - m2
    assert all(v == 0.0 for v in diff.atom_contacts.counter.values())
    assert len(diff.atom_contacts.counter) == len(atoms)
    with pytest.warns(PendingDeprecationWarning):
        cm.ContactFrequency.from_contacts(atoms, residues, 5, traj.topology)
    with pytest.raises(RuntimeError):
        cm.ContactMap()
    assert cm.residue_neighborhood(traj.topology.residue(0), n=2) == [0, 1, 2]
    assert cm.residue_neighborhood(traj.topology.residue(2), n=0) == [2]
    import pickle
    again = pickle.loads(pickle.dumps(m2))
    assert again == m2
    rt = cm.ContactFrequency.from_json(m2.to_json(), traj.topology)      # earlier two-argument form
    assert rt == m2
    assert cm.ContactFrequency.from_json(m2.to_json()) == m2              # the reference's signature


def test_json_files_written_by_the_reference_load_here():
    """tests/golden/ref_json.json holds the reference's own to_json / to_dict output (generated by
    make_golden_json.py, which also checks that the reference loads THIS package's JSON): same key
    set, same counters, one-argument from_json / from_dict (contact_map.py:202-333,838-880;
    contact_trajectory.py:152-162)."""
    import json
    with open(os.path.join(goldens.GOLDEN_DIR, "ref_json.json")) as f:
        gold = json.load(f)
    want_atoms = collections.Counter({frozenset(k): v for k, v in gold["expected"]["atom_contacts"]})
    want_res = collections.Counter({frozenset(k): v for k, v in gold["expected"]["residue_contacts"]})
    freq = cm.ContactFrequency.from_json(gold["frequency_json"])
    assert freq._atom_contacts == want_atoms and freq._residue_contacts == want_res
    assert freq.n_frames == gold["expected"]["n_frames"] and freq.cutoff == 0.075
    assert freq.topology.n_atoms == 10 and freq.topology.n_residues == 5
    again = json.loads(freq.to_json())
    ref_dct = json.loads(gold["frequency_json"])
    assert set(again) == set(ref_dct)
    for key in ("cutoff", "n_neighbors_ignored", "n_frames", "use_atom_slice"):
        assert again[key] == ref_dct[key]
    for key in ("query", "haystack", "query_res_idx", "haystack_res_idx", "all_atoms", "all_residues"):
        assert sorted(again[key]) == sorted(ref_dct[key])
    norm = lambda text: {tuple(sorted(json.loads(k))): v for k, v in json.loads(text).items()}   # noqa: E731
    assert norm(again["atom_contacts"]) == norm(ref_dct["atom_contacts"])
    assert norm(again["residue_contacts"]) == norm(ref_dct["residue_contacts"])
    ctraj = cm.ContactTrajectory.from_dict(gold["trajectory_dict"])
    assert len(ctraj) == 5
    assert ctraj.contact_frequency()._atom_contacts == want_atoms
    assert cm.ContactTrajectory.from_json(ctraj.to_json()) == ctraj
    diff = cm.ContactDifference.from_dict(gold["difference_dict"])
    want_diff = {frozenset(k): v for k, v in gold["expected"]["difference_atom"]}
    assert dict(diff.atom_contacts.counter) == want_diff
    assert set(diff.to_dict()) == set(gold["difference_dict"])


def test_query_haystack_containers_follow_the_reference():
    """query / haystack arrive in any order, with duplicates, as lists, sets, ranges or arrays; the
    containers the reference exposes (contact_map.py:132-139,409-422) come out the same whichever."""
    case = goldens.load("pdb_c0075_n0")
    top = goldens.trajectory(case).topology
    import warnings
    for query, haystack in (([7, 2, 2, 5], {9, 0, 4}), (np.array([5, 7, 2]), range(0, 10, 4)),
                            ((q for q in (2, 5, 7)), [4, 0, 9, 9])):
        if isinstance(haystack, range):
            haystack = [0, 4, 9]                                  # same set as the other spellings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = cm.ContactFrequency.from_contacts({}, {}, n_frames=1, topology=top, query=query,
                                                  haystack=haystack, cutoff=0.075, n_neighbors_ignored=0)
        assert m._query == {2, 5, 7} and m._haystack == {0, 4, 9}
        assert sorted(m.query) == [2, 5, 7] and sorted(m.haystack) == [0, 4, 9]
        assert m._all_atoms == (0, 2, 4, 5, 7, 9) and m.all_atoms == [0, 2, 4, 5, 7, 9]
        assert all(type(a) is int for a in m._all_atoms)
        assert m.query_range == (2, 8) and m.haystack_range == (0, 10)
        assert m._query_res_idx == {1, 2, 3} and m._haystack_res_idx == {0, 2, 4}
        assert m._all_residues == {0, 1, 2, 3, 4}
        assert m.use_atom_slice is True
    # defaults: both None share one selection
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        d = cm.ContactFrequency.from_contacts({}, {}, n_frames=1, topology=top, cutoff=0.075, n_neighbors_ignored=0)
    assert d._query == d._haystack == set(range(10)) and d.use_atom_slice is False


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from contact_map_b200.engine import EngineError
    case = goldens.load("pdb_c0075_n0")
    with pytest.raises(EngineError):
        cm.ContactFrequency(goldens.trajectory(case))
    with pytest.raises(EngineError):
        cm.MinimumDistanceCounter(goldens.trajectory(case), [0], [5])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "contact_map_b200")
    for dirpath, _, files in os.walk(pkg):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, name)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), name
                assert "cm_oracle" not in text, name


def test_cabi_exports_every_declared_symbol():
    from contact_map_b200 import _lib
    header = open(os.path.join(ROOT, "include", "cmb200.h")).read()
    declared = sorted(set(re.findall(r"\b(cmb_[a-z_0-9]+)\s*\(", header)))
    assert declared == sorted(_lib.EXPORTS)
    _lib.build()
    handle = ctypes.CDLL(_lib.SO_PATH)
    for name in declared:
        assert hasattr(handle, name), name
    assert handle.cmb_version() == 100


def test_synth_generator_properties():
    s1 = synth.kras_like()
    a = s1.frames(0, 3)
    b = s1.frames(2, 1)
    assert np.array_equal(a[2], b[0])                 # frames are pure functions of their index
    assert a.dtype == np.float32 and a.shape == (3, 2700, 3)
    # image shifts move whole residues by lattice vectors: per-residue offsets are multiples of the box
    plain = synth.make_frames(s1.base, s1.groups, s1.box, s1.seed, s1.amplitude, False, 0, 1)[0]
    shift = (a[0] - plain) / np.diag(s1.box)
    assert np.allclose(shift, np.round(shift), atol=1e-5)
    assert set(np.unique(np.round(shift)).tolist()) <= {-1.0, 0.0, 1.0}
    s3 = synth.solvated_triclinic(n_atoms=3000, lattice=15)
    assert s3.topology.n_chains == 2 and s3.box[1, 0] != 0


def test_trajectory_container(tmp_path):
    case = goldens.load("syn_ortho_sub")
    traj = goldens.trajectory(case)
    assert len(traj[1:3]) == 2 and len(traj[0]) == 1 and traj[-1].xyz.shape[0] == 1
    sub = traj.atom_slice([0, 1, 2, 9])
    assert sub.n_atoms == 4 and sub.topology.n_atoms == 4
    assert traj.topology == traj.topology.copy() and hash(traj.topology) == hash(traj.topology.copy())
    assert list(traj.topology.select("all")) == list(range(traj.n_atoms))
    pdb = tmp_path / "t.pdb"
    lines = ["CRYST1   25.000   25.000   25.000  90.00  90.00  90.00 P 1           1"]
    for i, (x, y) in enumerate([(0.25, 0.75), (0.25, 1.25), (2.25, 2.75)]):
        lines.append("HETATM%5d  C%d  XXX A%4d    %8.3f%8.3f%8.3f  1.00  0.00           C"
                     % (i + 1, i % 2 + 1, i // 2 + 1, x, y, 0.0))
    lines.append("ENDMDL")
    pdb.write_text("\n".join(lines) + "\n")
    t = cm.load(str(pdb))
    assert t.n_atoms == 3 and t.topology.n_residues == 2
    assert t.xyz[0, 2].tolist() == [np.float32(0.225), np.float32(0.275), 0.0]
    assert np.array_equal(t.unitcell_vectors[0], np.diag([2.5, 2.5, 2.5]).astype(np.float32))


def test_memory_mapped_npy_reader_is_lazy(tmp_path):
    """Chunked reader (reference load_trajectory_task, frequency_task.py:70-88): a .npy coordinate
    file is memory-mapped, slicing frames copies nothing, and every task sees exactly its frames."""
    from contact_map_b200 import frequency_task, trajectory
    case = goldens.load("syn_tric_all")
    traj = goldens.trajectory(case)
    name = str(tmp_path / "traj.npy")
    sidecar = trajectory.save_npy(name, traj)
    assert sidecar.endswith("traj.top.npz")
    lazy = cm.load(name)
    assert isinstance(lazy.xyz.base, np.memmap) or isinstance(lazy.xyz, np.memmap)
    assert not lazy.xyz.flags["OWNDATA"] and not lazy.xyz.flags["WRITEABLE"]
    assert lazy.topology == traj.topology
    assert np.array_equal(lazy.unitcell_vectors, traj.unitcell_vectors)
    assert np.array_equal(lazy.xyz, traj.xyz)
    # slices of the mapped file stay views of it (nothing is materialised)
    total = 0
    for sl in frequency_task.default_slices(len(traj), 3):
        part = frequency_task.load_trajectory_task(sl, name)
        assert np.shares_memory(part.xyz, lazy.xyz) or not part.xyz.flags["OWNDATA"]
        assert np.array_equal(part.xyz, traj.xyz[sl])
        assert np.array_equal(part.unitcell_vectors, traj.unitcell_vectors[sl])
        total += len(part)
    assert total == len(traj)
    # MDTraj's `top=` keyword: topology from another file / object; no side-car needed then
    bare = str(tmp_path / "bare.npy")
    np.save(bare, traj.xyz)
    assert cm.load(bare, top=traj).topology == traj.topology
    assert cm.load(bare, top=sidecar).n_atoms == traj.n_atoms
    with pytest.raises(IOError):
        np.save(str(tmp_path / "bad.npy"), traj.xyz.astype(np.float64))
        cm.load(str(tmp_path / "bad.npy"), top=traj)


def test_pair_table_widens_int32_arrays_lazily():
    """The GPU export hands PairTable int32 index / count arrays; every reader sees int64 (no int32
    overflow in packed keys), the stored arrays are only widened when they are first read."""
    from contact_map_b200.contact_count import PairTable
    i = np.array([0, 70000, 99998], dtype=np.int32)
    j = np.array([5, 99999, 99999], dtype=np.int32)
    v = np.array([3, 1, 4096], dtype=np.uint32).view(np.int32)
    t = PairTable(i, j, v)
    assert len(t) == 3 and t._i.dtype == np.int32 and t._value.dtype == np.int32      # untouched so far
    assert t.i.dtype == np.int64 and t.j.dtype == np.int64 and t.value.dtype == np.int64
    assert t._i.dtype == np.int64                                                     # widened once, kept
    packed = t._packed(100000)
    assert packed.dtype == np.int64 and packed[1] == 70000 * 100000 + 99999            # would overflow int32
    both = t.combine(PairTable(np.array([0]), np.array([5]), np.array([2])))
    assert both.to_counter() == {frozenset((0, 5)): 5, frozenset((70000, 99999)): 1, frozenset((99998, 99999)): 4096}
    minus = t.combine(PairTable(np.array([0]), np.array([5]), np.array([3])), sign=-1)
    assert frozenset((0, 5)) not in minus.to_counter()
    assert t.scaled(4).value.dtype == np.float64 and t.scaled(4).value[0] == 0.75
    assert t.restricted(np.array([0, 5])).to_counter() == {frozenset((0, 5)): 3}
    # other dtypes are normalised at construction, floats stay floats
    u = PairTable(np.array([1], dtype=np.uint8), [2], np.array([0.5]))
    assert u.i.dtype == np.int64 and u.j.dtype == np.int64 and u.value.dtype == np.float64
