"""GPU parity: engine (C ABI) vs the CPU oracle on seeded random systems."""
import numpy as np
import pytest
import torch

from contact_map_b200 import synth
from oracle import clib
from tests.helpers import system_arrays, dense_from_export, keys_to_frames

pytestmark = pytest.mark.gpu

CASES = [
    # n_atoms, box, cutoff, n_ignored, n_frames, q_frac, h_frac
    (60, "ortho", 0.45, 2, 6, 1.0, 1.0),
    (257, "ortho", 0.45, 0, 5, 1.0, 1.0),
    (300, "triclinic", 0.45, 2, 5, 1.0, 1.0),
    (300, "none", 0.45, 1, 5, 1.0, 1.0),
    (500, "ortho", 0.30, 3, 4, 0.3, 0.6),
    (500, "triclinic", 0.55, 2, 4, 0.5, 0.5),
    (411, "none", 0.60, 2, 4, 0.2, 1.0),
    (1500, "ortho", 0.45, 2, 3, 1.0, 1.0),
    (10, "ortho", 0.45, 2, 5, 1.0, 1.0),
    (3, "none", 0.45, 0, 2, 1.0, 1.0),
    # too large for two CTAs per SM: the 1024-thread instantiation of the fused kernel
    (3500, "triclinic", 0.45, 2, 2, 1.0, 1.0),
    (4100, "none", 0.50, 1, 2, 0.6, 0.8),
    (6144, "ortho", 0.45, 2, 2, 1.0, 1.0),
]


def _run_case(engine, case, seed):
    n, box_kind, cutoff, n_ign, n_frames, qf, hf = case
    rng = np.random.default_rng(seed)
    extent = max(1.2, (n / 50.0) ** (1.0 / 3.0))
    system = synth.random_system(rng, n, n_chains=3, box_kind=box_kind, extent=extent)
    query = np.sort(rng.choice(n, size=max(1, int(qf * n)), replace=False))
    hay = np.sort(rng.choice(n, size=max(1, int(hf * n)), replace=False))
    atom_res, atom_chain, role = system_arrays(system, query, hay)
    xyz = system.frames(0, n_frames)
    box = None if system.box is None else np.broadcast_to(system.box, (n_frames, 3, 3)).copy()
    n_res = system.topology.n_residues
    ref_atom, ref_res = clib.accumulate_dense(xyz, box, cutoff, atom_res, atom_chain, role, n_ign, n_res)
    return system, xyz, box, (atom_res, atom_chain, role), (ref_atom, ref_res)


@pytest.mark.parametrize("case_id", range(len(CASES)))
def test_accumulate_matches_oracle(engine, case_id):
    case = CASES[case_id]
    n, _, cutoff, n_ign, n_frames = case[0], case[1], case[2], case[3], case[4]
    system, xyz, box, (atom_res, atom_chain, role), (ref_atom, ref_res) = _run_case(engine, case, 100 + case_id)
    engine.set_system(atom_res, atom_chain, role, cutoff, n_ign)
    atom_t, res_t = engine.new_tables()
    d_xyz = torch.as_tensor(xyz, device=engine.device)
    d_box = None if box is None else torch.as_tensor(box, device=engine.device)
    engine.accumulate(d_xyz, d_box, atom_t, res_t)
    idx, cnt = engine.export_nonzero(atom_t)
    got_atom = dense_from_export(idx, cnt, n)
    assert np.array_equal(got_atom, ref_atom), "atom counts differ (%d vs %d pairs)" % (
        np.count_nonzero(got_atom), np.count_nonzero(ref_atom))
    ridx, rcnt = engine.export_nonzero(res_t)
    rc = engine.n_res_c
    got_res_c = dense_from_export(ridx, rcnt, rc)
    rmap = engine.residue_map
    ref_c = ref_res[np.ix_(rmap, rmap)]
    assert np.array_equal(got_res_c, ref_c)
    # nothing outside the used residues in the oracle
    assert ref_res.sum() == ref_c.sum()
    assert ref_atom.sum() > 0 or n < 20


@pytest.mark.parametrize("case_id", [0, 2, 3, 4])
def test_host_path_and_frame_keys(engine, case_id):
    case = CASES[case_id]
    n, _, cutoff, n_ign, n_frames = case[0], case[1], case[2], case[3], case[4]
    system, xyz, box, (atom_res, atom_chain, role), (ref_atom, ref_res) = _run_case(engine, case, 100 + case_id)
    engine.set_system(atom_res, atom_chain, role, cutoff, n_ign)
    # host-buffer entry point, tiny chunks to exercise the double buffering: pageable memory
    # (host threads stage it into pinned buffers), pinned memory (direct async copies), and
    # pageable memory with the staging switched off
    pinned = torch.empty(xyz.shape, dtype=torch.float32, pin_memory=True)
    pinned.copy_(torch.as_tensor(xyz))
    for host_xyz, staged in ((xyz, 1), (pinned.numpy(), 1), (xyz, 0)):
        engine.set_option("stage_pageable", staged)
        atom_t, res_t = engine.new_tables()
        engine.accumulate_host(host_xyz, box, atom_t, res_t, chunk_frames=2)
        idx, cnt = engine.export_nonzero(atom_t)
        assert np.array_equal(dense_from_export(idx, cnt, n), ref_atom)
    engine.set_option("stage_pageable", 1)
    # per-frame keys
    d_xyz = torch.as_tensor(xyz, device=engine.device)
    d_box = None if box is None else torch.as_tensor(box, device=engine.device)
    akeys, rkeys = engine.frame_contacts(d_xyz, d_box, frame0=0)
    frame, lo, hi = keys_to_frames(akeys)
    n_res = system.topology.n_residues
    for f in range(n_frames):
        ref_pairs = clib.contacts_frame(xyz[f], None if box is None else box[f], cutoff,
                                        atom_res, atom_chain, role, n_ign)
        sel = frame == f
        got = np.stack([lo[sel], hi[sel]], axis=1)
        assert np.array_equal(got, ref_pairs), "frame %d" % f
    assert np.all(np.diff(akeys.astype(np.int64)) > 0)
    # residue keys agree with the per-frame residue sets derived from the atom pairs
    rframe, rlo, rhi = keys_to_frames(rkeys)
    rmap = engine.residue_map
    for f in range(n_frames):
        sel = frame == f
        ra, rb = atom_res[lo[sel]], atom_res[hi[sel]]
        want = set(zip(np.minimum(ra, rb).tolist(), np.maximum(ra, rb).tolist()))
        rs = rframe == f
        got = set(zip(rmap[rlo[rs]].tolist(), rmap[rhi[rs]].tolist()))
        assert got == want


@pytest.mark.parametrize("case_id,n_threads", [(0, 1), (2, 4), (7, 0)])
def test_host_gather_path(engine, case_id, n_threads):
    """cmb_accumulate_host_gather (fused _atom_slice, atom_indexer.py:5-22): a host trajectory that
    holds extra atoms gives the same tables as the pre-sliced one."""
    case = CASES[case_id]
    n, cutoff, n_ign, n_frames = case[0], case[2], case[3], case[4]
    system, xyz, box, (atom_res, atom_chain, role), (ref_atom, _) = _run_case(engine, case, 100 + case_id)
    rng = np.random.default_rng(7 + case_id)
    n_total = n + 37
    used = np.sort(rng.choice(n_total, size=n, replace=False)).astype(np.int32)
    full = rng.normal(size=(n_frames, n_total, 3)).astype(np.float32)
    full[:, used] = xyz
    engine.set_system(atom_res, atom_chain, role, cutoff, n_ign)
    atom_t, res_t = engine.new_tables()
    engine.accumulate_host_gather(full, used, box, atom_t, res_t, chunk_frames=2, n_threads=n_threads)
    idx, cnt = engine.export_nonzero(atom_t)
    assert np.array_equal(dense_from_export(idx, cnt, n), ref_atom)
    with pytest.raises(Exception):
        engine.accumulate_host_gather(full, used[::-1].copy(), box, atom_t, res_t)


def test_export_sorted_and_wide_counts(engine):
    """cmb_export_sorted (packed, device-sorted) and the unpacked fallback for counts >= 2^24."""
    rng = np.random.default_rng(3)
    table = np.zeros(5003, dtype=np.uint32)
    where = np.sort(rng.choice(5003, size=700, replace=False))
    table[where] = rng.integers(1, 1 << 20, size=700)
    t = torch.as_tensor(table.view(np.int32), device=engine.device)
    idx, cnt = engine.export_nonzero(t)
    assert np.array_equal(idx, where) and np.array_equal(cnt, table[where])
    table[where[5]] = (1 << 24) + 7                        # does not fit the packed form
    table[5002] = 0xffffffff
    t = torch.as_tensor(table.view(np.int32), device=engine.device)
    idx, cnt = engine.export_nonzero(t)
    nz = np.flatnonzero(table)
    assert np.array_equal(idx, nz) and np.array_equal(cnt, table[nz])
    empty = torch.zeros(64, dtype=torch.int32, device=engine.device)
    idx, cnt = engine.export_nonzero(empty)
    assert len(idx) == 0 and len(cnt) == 0


@pytest.mark.parametrize("box_kind", ["none", "ortho", "triclinic"])
def test_crowded_cells(engine, box_kind):
    """Cells holding far more than 32 atoms: own atoms are screened in sub-batches of 32, the
    candidate set spans several 128-slot chunks, and the ring fills faster than it drains (the
    slot-by-slot push path).  350 atoms inside a 0.9 nm blob, optionally in a periodic box."""
    rng = np.random.default_rng(99)
    n, n_frames, cutoff = 350, 3, 0.45
    blob = rng.uniform(0.0, 0.9, size=(n, 3))
    if box_kind == "none":
        box = None
    elif box_kind == "ortho":
        box = np.diag([2.9, 3.3, 2.5]).astype(np.float32)
    else:
        box = np.array([[2.9, 0, 0], [0.7, 3.3, 0], [-0.5, 0.9, 2.5]], dtype=np.float32)
    xyz = np.empty((n_frames, n, 3), dtype=np.float32)
    for f in range(n_frames):
        p = blob + rng.normal(scale=0.02, size=(n, 3)) + np.array([2.7, 0.1, 1.2])    # straddles box faces
        if box is not None:
            p = p + rng.integers(-1, 2, size=(n, 3)).astype(np.float64) @ box.astype(np.float64)
        xyz[f] = p.astype(np.float32)
    res = (np.arange(n) // 5).astype(np.int32)
    chain = (res // 20).astype(np.int32)
    role = np.full(n, 3, dtype=np.uint8)
    boxes = None if box is None else np.broadcast_to(box, (n_frames, 3, 3)).copy()
    n_res = int(res.max()) + 1
    ref_atom, ref_res = clib.accumulate_dense(xyz, boxes, cutoff, res, chain, role, 1, n_res)
    assert ref_atom.sum() > 20000
    for force in (0, 1):
        engine.set_option("force_path", force if force else -1)
        try:
            engine.set_system(res, chain, role, cutoff, 1)
            atom_t, res_t = engine.new_tables(hashed=False)
            engine.accumulate(torch.as_tensor(xyz, device=engine.device),
                              None if boxes is None else torch.as_tensor(boxes, device=engine.device),
                              atom_t, res_t)
            (idx, cnt), (ridx, rcnt) = engine.export_many([atom_t, res_t])
            assert np.array_equal(dense_from_export(idx, cnt, n), ref_atom)
            rmap = engine.residue_map
            assert np.array_equal(dense_from_export(ridx, rcnt, engine.n_res_c), ref_res[np.ix_(rmap, rmap)])
        finally:
            engine.set_option("force_path", -1)


@pytest.mark.parametrize("case_id", [0, 4, 7])
def test_export_pair_views_equals_export_many(engine, case_id):
    """cmb_unpack_pairs (row / column / count triples split on the device, one scan over both
    tables) against the packed export split on the host; empty tables, a capacity guess that is too
    small (exact retry) and counts beyond 24 bits (refused: callers fall back) included."""
    case = CASES[case_id]
    system, xyz, box, (atom_res, atom_chain, role), _ = _run_case(engine, case, 300 + case_id)
    engine.set_system(atom_res, atom_chain, role, case[2], case[3])
    n, rc = engine.n_used, engine.n_res_c
    atom_t, res_t = engine.new_tables()
    (lo, hi, cnt), (rlo, rhi, rcnt) = engine.export_pair_views(atom_t, res_t)
    assert len(lo) == len(rlo) == 0
    d_xyz = torch.as_tensor(xyz, device=engine.device)
    d_box = None if box is None else torch.as_tensor(box, device=engine.device)
    for rounds in (1, 3):                      # the second pass quadruples the entry count hint's tables
        for _ in range(rounds):
            engine.accumulate(d_xyz, d_box, atom_t, res_t)
        (idx, c), (ridx, rcn) = engine.export_many([atom_t, res_t])
        (lo, hi, cnt), (rlo, rhi, rcnt) = engine.export_pair_views(atom_t, res_t)
        assert np.array_equal(lo.astype(np.int64) * n + hi, idx) and np.array_equal(cnt, c)
        assert np.array_equal(rlo.astype(np.int64) * rc + rhi, ridx) and np.array_equal(rcnt, rcn)
        assert lo.dtype == np.int32 and cnt.dtype == np.uint32 and len(idx) > 0
    # a stale, too small capacity hint: one exact retry
    engine._export_hint.clear()
    engine._export_pool.clear()
    small_a, small_r = engine.new_tables()
    small_a[1] = 5
    engine.export_pair_views(small_a, small_r)              # hint = 1 entry
    engine._export_pool.clear()
    (lo2, hi2, cnt2), (rlo2, _, _) = engine.export_pair_views(atom_t, res_t)
    assert np.array_equal(lo2.astype(np.int64) * n + hi2, idx) and np.array_equal(cnt2, c) and len(rlo2) == len(ridx)
    # counts of 2^24 and more do not fit the packed words
    small_a[1] = 1 << 24
    assert engine.export_pair_views(small_a, small_r) is None
    # tables that are not the two views of one allocation
    assert engine.export_pair_views(atom_t.clone(), res_t) is None
