"""GPU: the public Python API against the golden vectors of the UNMODIFIED reference.

Assertions mirror the reference's own tests (test_contact_map.py, test_contact_trajectory.py,
test_min_dist.py, test_frequency_task.py), which compare the private integer Counters and the
public frequency Counters directly.
"""
import collections
import warnings

import numpy as np
import pytest

import contact_map_b200 as cm
from tests import goldens

pytestmark = pytest.mark.gpu
warnings.simplefilter("ignore", PendingDeprecationWarning)


@pytest.mark.parametrize("name", goldens.CASE_NAMES)
def test_contact_frequency_counters(engine, name):
    case = goldens.load(name)
    traj = goldens.trajectory(case)
    m = cm.ContactFrequency(traj, **goldens.call_kwargs(case))
    assert m._atom_contacts == goldens.counter(case["atom_pairs"], case["atom_counts"])
    assert m._residue_contacts == goldens.counter(case["res_pairs"], case["res_counts"])
    assert m.n_frames == len(traj)
    assert m.use_atom_slice == bool(case["use_atom_slice"])
    assert sorted(m.query) == case["query"].tolist()
    assert sorted(m.haystack) == case["haystack"].tolist()
    ac = m.atom_contacts
    assert ac.counter == goldens.counter(case["atom_freq_pairs"], case["atom_freq_values"])
    assert ac.n_x == tuple(case["n_x"].tolist()) and ac.n_y == tuple(case["n_y"].tolist())
    assert ac.max_size == int(case["max_size"])
    rc = m.residue_contacts
    assert rc.n_x == tuple(case["res_n_x"].tolist()) and rc.n_y == tuple(case["res_n_y"].tolist())
    assert rc.max_size == int(case["res_max_size"])
    # keys are frozensets of plain Python ints, values plain ints
    for key, value in m._atom_contacts.items():
        assert all(type(k) is int for k in key) and type(value) is int
    # dense matrix export agrees with the counter
    dense = ac.sparse_matrix.toarray()
    assert dense.shape == (ac.max_size, ac.max_size)
    for key, value in ac.counter.items():
        i, j = tuple(key)
        assert dense[i, j] == value and dense[j, i] == value
    assert np.count_nonzero(dense) == 2 * len(ac.counter)


@pytest.mark.parametrize("name", goldens.CASE_NAMES)
def test_contact_trajectory_frames(engine, name):
    case = goldens.load(name)
    traj = goldens.trajectory(case)
    ct = cm.ContactTrajectory(traj, **goldens.call_kwargs(case))
    assert len(ct) == len(traj)
    want_atoms = goldens.frame_sets(case["frame_atom_ptr"], case["frame_atom_pairs"])
    want_res = goldens.frame_sets(case["frame_res_ptr"], case["frame_res_pairs"])
    for f, frame in enumerate(ct):
        assert frame.n_frames == 1
        assert frame._atom_contacts == collections.Counter(want_atoms[f])
        assert frame._residue_contacts == collections.Counter(want_res[f])
    assert [set(c.counter) for c in ct.atom_contacts] == want_atoms
    freq = ct.contact_frequency()
    assert freq._atom_contacts == goldens.counter(case["atom_pairs"], case["atom_counts"])
    assert freq._residue_contacts == goldens.counter(case["res_pairs"], case["res_counts"])
    assert freq.n_frames == len(traj)
    # slicing, negative indices, joins (reference contact_trajectory.py:91-95,185-200)
    assert ct[-1]._atom_contacts == collections.Counter(want_atoms[-1])
    assert [m._atom_contacts for m in ct[1:3]] == [collections.Counter(s) for s in want_atoms[1:3]]
    if len(traj) >= 3:
        first = cm.ContactTrajectory(traj[:2], **goldens.call_kwargs(case))
        second = cm.ContactTrajectory(traj[2:], **goldens.call_kwargs(case))
        joined = cm.ContactTrajectory.join([first, second])
        assert joined == ct and hash(joined) == hash(ct)


def test_accumulate_and_subtract(engine):
    """add/subtract_contact_frequency (reference test_contact_map.py:508-543)."""
    case = goldens.load("pdb_c0075_n0")
    traj = goldens.trajectory(case)
    kw = goldens.call_kwargs(case)
    total = cm.ContactFrequency(traj, **kw)
    head = cm.ContactFrequency(traj[:3], **kw)
    tail = cm.ContactFrequency(traj[3:], **kw)
    head.add_contact_frequency(tail)
    assert head == total and head.n_frames == 5
    assert hash(head) == hash(total)
    total.subtract_contact_frequency(tail)
    again = cm.ContactFrequency(traj[:3], **kw)
    assert total._atom_contacts == again._atom_contacts
    assert total._residue_contacts == again._residue_contacts
    assert total.n_frames == 3
    other = cm.ContactFrequency(traj, cutoff=0.45, n_neighbors_ignored=0)
    with pytest.raises(AssertionError):
        total.add_contact_frequency(other)


def test_contact_difference(engine):
    case = goldens.load("pdb_c0075_n0")
    diffs = goldens.load("pdb_differences")
    traj = goldens.trajectory(case)
    kw = goldens.call_kwargs(case)
    f_all = cm.ContactFrequency(traj, **kw)
    f0 = cm.ContactFrequency(traj[0], **kw)
    f4 = cm.ContactFrequency(traj[4], **kw)
    for tag, diff in (("traj_minus_f0", f_all - f0), ("f4_minus_f0", f4 - f0), ("f0_minus_traj", f0 - f_all)):
        assert isinstance(diff, cm.ContactDifference)
        assert diff.atom_contacts.counter == goldens.counter(diffs[tag + "_atom_pairs"], diffs[tag + "_atom_values"])
        assert diff.residue_contacts.counter == goldens.counter(diffs[tag + "_res_pairs"], diffs[tag + "_res_values"])
    # literal values of the reference test (test_contact_map.py:591-609): zeros stay in the Counter
    d = (f_all - f0).atom_contacts.counter
    assert d[frozenset([1, 4])] == 0.8 - 1.0 and d[frozenset([4, 6])] == 0.0
    assert d[frozenset([0, 9])] == 0.2
    # mismatched parameters become None (fix_parameters.py:34-36)
    other = cm.ContactFrequency(traj[0], cutoff=0.1, n_neighbors_ignored=1)
    mixed = f_all - other
    assert mixed.cutoff is None and mixed.n_neighbors_ignored is None
    with pytest.raises(RuntimeError):
        f_all.contacts["bad"]
    with pytest.raises(NotImplementedError):
        (f_all - f0) - f0


def test_rolling_frequency(engine):
    case = goldens.load("pdb_c0075_n0")
    gold = goldens.load("pdb_rolling")
    ct = cm.ContactTrajectory(goldens.trajectory(case), **goldens.call_kwargs(case))
    for width, step in ((2, 1), (3, 2), (5, 1)):
        windows = list(cm.RollingContactFrequency(ct, width=width, step=step))
        assert len(windows) == int(gold["w%d_s%d_n" % (width, step)])
        for w, cmap in enumerate(windows):
            tag = "w%d_s%d_%d" % (width, step, w)
            assert cmap._atom_contacts == goldens.counter(gold[tag + "_pairs"], gold[tag + "_counts"])
            assert cmap.n_frames == int(gold[tag + "_nframes"])
    assert len(list(ct.rolling_frequency(window_size=2, step=1))) == 4


@pytest.mark.parametrize("name", ["pdb_c0075_n0", "syn_ortho_all", "syn_tric_all", "syn_open_all"])
def test_min_distance_counter(engine, name):
    case = goldens.load(name)
    traj = goldens.trajectory(case)
    q, h = case["mindist_query"].tolist(), case["mindist_haystack"].tolist()
    mdc = cm.MinimumDistanceCounter(traj, q, h)
    # north_star: within 1e-6 nm; the implementation is bit-equal
    assert np.array_equal(np.asarray(mdc.minimum_distances), case["mindist_values"])
    assert np.array_equal(np.asarray(mdc._min_pairs), case["mindist_pairs"])
    assert mdc.atom_pairs[:2] == [(q[0], h[0]), (q[0], h[1])]
    assert len(mdc.atom_pairs) == len(q) * len(h)
    if "mindist_history" in case:
        assert [[a.index, b.index] for a, b in mdc.atom_history] == case["mindist_history"].tolist()
        assert mdc.atom_count[(traj.topology.atom(5), traj.topology.atom(6))] == 2
        assert sum(mdc.residue_count.values()) == len(traj)


def test_nearest_atoms(engine):
    case = goldens.load("pdb_c0075_n0")
    traj = goldens.trajectory(case)
    for frame, excluded, tag in ((0, None, "f0_res"), (4, None, "f4_res"), (4, {}, "f4_all")):
        na = cm.NearestAtoms(traj, cutoff=0.075, frame_number=frame, excluded=excluded)
        atoms = sorted(na.nearest)
        assert atoms == case["nearest_%s_atom" % tag].tolist()
        assert [na.nearest[a] for a in atoms] == case["nearest_%s_partner" % tag].tolist()
        got = np.array([na.nearest_distance[a] for a in atoms], dtype=np.float32)
        assert np.array_equal(got, case["nearest_%s_dist" % tag])
        dists = [d for _, _, d in na.sorted_distances]
        assert dists == sorted(dists)
    # explicit exclusion dict (CSR path) equal to the same-residue default
    atom_res = case["atom_residue"]
    explicit = {i: np.nonzero(atom_res == atom_res[i])[0].tolist() for i in range(len(atom_res))}
    a = cm.NearestAtoms(traj, cutoff=0.075, frame_number=4, excluded=explicit)
    b = cm.NearestAtoms(traj, cutoff=0.075, frame_number=4)
    assert a.nearest == b.nearest and a.nearest_distance == b.nearest_distance
    for name in ("syn_ortho_all", "syn_tric_all", "syn_open_all"):
        case = goldens.load(name)
        na = cm.NearestAtoms(goldens.trajectory(case), cutoff=float(case["cutoff"]), frame_number=1)
        atoms = sorted(na.nearest)
        assert atoms == case["nearest_atom"].tolist()
        assert [na.nearest[a] for a in atoms] == case["nearest_partner"].tolist()
        got = np.array([na.nearest_distance[a] for a in atoms], dtype=np.float32)
        assert np.array_equal(got, case["nearest_dist"])


def test_chunked_runner_equals_serial(engine, tmp_path):
    """The reference's own check (test_dask_runner.py:74-90): runner == serial, default params."""
    case = goldens.load("syn_ortho_all")
    traj = goldens.trajectory(case)
    serial = cm.ContactFrequency(traj)
    for n_blocks in (1, 2, 3):
        chunked = cm.ChunkedContactFrequency(traj, n_blocks=n_blocks)
        assert chunked._atom_contacts == serial._atom_contacts
        assert chunked._residue_contacts == serial._residue_contacts
        assert chunked.n_frames == serial.n_frames
    path = str(tmp_path / "traj.npz")
    np.savez(path, xyz=traj.xyz, unitcell_vectors=traj.unitcell_vectors,
             atom_residue=traj.topology.atom_residue, residue_chain=traj.topology.residue_chain,
             atom_elements=np.array(traj.topology.atom_elements),
             residue_names=np.array(traj.topology.residue_names))
    dask_like = cm.DaskContactFrequency(client=3, filename=path)
    assert dask_like._atom_contacts == serial._atom_contacts
    assert dask_like.run_info["trajectory_file"] == path
    dtraj = cm.DaskContactTrajectory(client=2, filename=path)
    assert dtraj == cm.ContactTrajectory(traj)
    # map/reduce task functions (test_frequency_task.py:51-91)
    ft = cm.frequency_task
    slices = ft.default_slices(len(traj), 3)
    maps = [ft.map_task(ft.load_trajectory_task(s, path), {"cutoff": 0.45, "n_neighbors_ignored": 2})
            for s in slices]
    reduced = ft.reduce_all_results(maps)
    assert reduced == serial


def test_device_resident_input_and_slicing(engine):
    import torch
    case = goldens.load("syn_tric_sub")
    traj = goldens.trajectory(case)
    kw = goldens.call_kwargs(case)
    host = cm.ContactFrequency(traj, **kw)

    class DeviceTraj(object):
        def __init__(self, t):
            self.topology = t.topology
            self.xyz = torch.as_tensor(t.xyz, device=engine.device)
            self.unitcell_vectors = torch.as_tensor(t.unitcell_vectors, device=engine.device)

        def __len__(self):
            return self.xyz.shape[0]

    dev = cm.ContactFrequency(DeviceTraj(traj), **kw)
    assert dev._atom_contacts == host._atom_contacts
    assert dev._residue_contacts == host._residue_contacts
    assert host._atom_contacts == goldens.counter(case["atom_pairs"], case["atom_counts"])


def test_public_api_on_hashed_tables(engine, monkeypatch):
    """Systems whose dense count table would be too large run on hashed tables transparently:
    force that regime (global-memory cell path + a tiny dense limit) and compare the public
    ContactFrequency / chunked runner results with the default (dense, fused) ones."""
    from contact_map_b200.engine import ContactEngine
    case = goldens.load("syn_tric_sub")
    traj = goldens.trajectory(case)
    kw = goldens.call_kwargs(case)
    want = cm.ContactFrequency(traj, **kw)
    engine.set_option("force_path", 1)
    monkeypatch.setattr(ContactEngine, "DENSE_LIMIT_BYTES", 1024)
    try:
        got = cm.ContactFrequency(traj, **kw)
        assert got._table_kind == "hashed"
        assert got._atom_contacts == want._atom_contacts == goldens.counter(case["atom_pairs"], case["atom_counts"])
        assert got._residue_contacts == want._residue_contacts
        chunked = cm.ChunkedContactFrequency(traj, n_blocks=3, **kw)
        assert chunked._atom_contacts == want._atom_contacts
        assert chunked._residue_contacts == want._residue_contacts
        assert chunked.n_frames == len(traj)
    finally:
        engine.set_option("force_path", -1)


def test_chunked_reader_streams_a_mapped_file(engine, tmp_path):
    """f2 (reference frequency_task.py:70-88, dask_runner.py:109,179): a memory-mapped trajectory
    file several times larger than the engine's staging buffers goes through the runner block by
    block -- plain copy (all atoms used) and fused gather (query/haystack subsets) -- and equals the
    in-memory run; the map/reduce tasks see only their own slices."""
    from contact_map_b200 import synth, trajectory
    system = synth.kras_like()
    n_frames = 1500                                        # 48.6 MB of coordinates
    xyz = system.frames(0, n_frames)
    box = np.broadcast_to(system.box, (n_frames, 3, 3)).copy()
    mem = cm.Trajectory(xyz, system.topology, box)
    name = str(tmp_path / "big.npy")
    trajectory.save_npy(name, mem)
    lazy = cm.load(name)
    assert not lazy.xyz.flags["OWNDATA"]
    engine.set_option("stage_chunk_frames", 100)           # staging buffers of 100 frames (3.2 MB)
    try:
        for kw in ({}, {"query": list(range(0, 300)), "haystack": list(range(200, 2700, 2))}):
            serial = cm.ContactFrequency(mem, **kw)
            from_file = cm.DaskContactFrequency(client=4, filename=name, **kw)
            assert from_file._atom_contacts == serial._atom_contacts
            assert from_file._residue_contacts == serial._residue_contacts
            assert from_file.n_frames == n_frames
            direct = cm.ContactFrequency(lazy, **kw)       # the plain class on the mapped file
            assert direct == serial
        ft = cm.frequency_task
        parts = [ft.map_task(ft.load_trajectory_task(s, name), {"cutoff": 0.45, "n_neighbors_ignored": 2})
                 for s in ft.default_slices(n_frames, 2)]
        assert ft.reduce_all_results(parts) == cm.ContactFrequency(mem)
    finally:
        engine.set_option("stage_chunk_frames", 0)
    assert len(serial._atom_contacts) > 100
