"""GPU: the frame-batch pair-list path (csrc/cmb_verlet.cuh) against the oracle and the cell
kernels -- list valid, guard trips, list unusable (automatic fall back), audit mode -- and the
100 000-atom parity checks of SURVEY 8(d) (S3 frames {0, 1, F/2, F-1} and a prefix; S4 min-dist /
nearest on the 100k system)."""
import numpy as np
import pytest
import torch

import contact_map_b200 as cm
from contact_map_b200 import synth
from oracle import clib, refpath
from tests.helpers import system_arrays, dense_from_export, keys_to_frames

pytestmark = pytest.mark.gpu

VL_STATS = ("vl_used", "vl_valid", "vl_fail", "vl_tripped", "vl_entries", "vl_tiles")


def _dev(engine, system, frame0, n_frames):
    dev = engine.device
    xyz = engine.synth_frames(torch.as_tensor(system.base, device=dev),
                              torch.as_tensor(system.groups.astype(np.int32), device=dev),
                              None if system.box is None else torch.as_tensor(system.box.reshape(9), device=dev),
                              system.seed, system.amplitude, system.image_shifts and system.box is not None,
                              frame0, n_frames)
    box = None if system.box is None else \
        torch.as_tensor(system.box, device=dev).reshape(1, 3, 3).expand(n_frames, 3, 3).contiguous()
    return xyz, box


def _accumulate(engine, xyz, box, mode, audit=0):
    engine.set_option("verlet", mode)
    engine.set_option("verlet_audit", audit)
    engine.set_option("verlet_max_rlist_pct", 200)      # the jittery random test systems need wide skins
    try:
        atom_t, res_t = engine.new_tables()
        engine.accumulate(xyz, box, atom_t, res_t)
        torch.cuda.synchronize()
        stats = {k: engine.get_stat(k) for k in VL_STATS}
    finally:
        engine.set_option("verlet", 1)
        engine.set_option("verlet_audit", 0)
        engine.set_option("verlet_max_rlist_pct", 175)
    return atom_t, res_t, stats


def _oracle_tables(engine, system, role, xyz, box, n_frames, cutoff, n_ignored):
    res, chain = system.topology.atom_residue_chain()
    xh = xyz[:n_frames].cpu().numpy()
    bh = None if box is None else box[:n_frames].cpu().numpy()
    want_a, want_r = clib.accumulate_dense(xh, bh, cutoff, res, chain, role, n_ignored, system.topology.n_residues)
    rmap = engine.residue_map
    return want_a, want_r[np.ix_(rmap, rmap)]


def _tables_np(engine, atom_t, res_t):
    n, rc = engine.n_used, engine.n_res_c
    return (atom_t.cpu().numpy().view(np.uint32).reshape(n, n), res_t.cpu().numpy().view(np.uint32).reshape(rc, rc))


@pytest.mark.parametrize("case", ["s1", "ligand", "ortho", "triclinic", "open", "n0", "tiles5000"])
def test_pair_list_path_matches_oracle(engine, case):
    """List valid, no trips: bit-equal to the oracle's dense tables (and to the cell kernels on
    the whole batch); the audit mode finds no pair where the screen and MDTraj's arithmetic
    disagree."""
    rng = np.random.default_rng(31)
    n_ignored, query, hay = 2, None, None
    if case == "s1":
        system, frames, check = synth.kras_like(), 300, 12
    elif case == "ligand":
        system, frames, check = synth.binding_system(ligand=True), 300, 12
        query, hay = system.query, system.haystack
    elif case == "tiles5000":
        # fused-kernel size (path 0) but more atoms than one tile holds: several tiles, guard pass
        system, frames, check = synth.solvated_triclinic(n_atoms=5000, lattice=18), 260, 6
    elif case == "n0":
        system, frames, check, n_ignored = synth.random_system(rng, 800, n_chains=2, box_kind="ortho", extent=2.6), 260, 16, 0
    else:
        system, frames, check = synth.random_system(rng, 1000, n_chains=3, box_kind=case, extent=2.7), 260, 16
        query = np.sort(rng.choice(1000, size=700, replace=False))
        hay = np.sort(rng.choice(1000, size=800, replace=False))
        n_ignored = 3
    res, chain, role = system_arrays(system, query, hay)
    engine.set_system(res, chain, role, 0.45, n_ignored)
    assert engine.path_kind == 0
    xyz, box = _dev(engine, system, 0, frames)
    cells_a, cells_r, st0 = _accumulate(engine, xyz, box, 0)
    assert st0["vl_used"] == 0
    list_a, list_r, st = _accumulate(engine, xyz, box, 2)
    assert st["vl_used"] == 1 and st["vl_valid"] == 1 and st["vl_tripped"] == 0 and st["vl_entries"] > 0
    assert (st["vl_tiles"] > 1) == (case == "tiles5000")
    assert torch.equal(cells_a, list_a) and torch.equal(cells_r, list_r)
    # a short batch of its own (its own list) against the oracle
    xs, bs = xyz[:check].contiguous(), None if box is None else box[:check].contiguous()
    got_a, got_r, st = _accumulate(engine, xs, bs, 2, audit=1)
    assert st["vl_valid"] == 1 and st["vl_tripped"] == 0
    assert engine.get_stat("vl_audited") > 0 and engine.get_stat("vl_disagree") == 0
    want_a, want_r = _oracle_tables(engine, system, role, xs, bs, check, 0.45, n_ignored)
    a, r = _tables_np(engine, got_a, got_r)
    assert np.array_equal(a, want_a) and np.array_equal(r, want_r)


def test_guard_trips_fall_back_to_the_cell_kernels(engine):
    """Frames the list cannot take -- atoms kicked beyond the skin mid-batch, a scrambled frame, a
    NaN coordinate, a frame with another unit cell -- land on the device-side fall-back list and are
    processed by the cell kernels: tables identical to a run without the list and to the oracle."""
    system = synth.kras_like()
    res, chain, role = system_arrays(system)
    engine.set_system(res, chain, role, 0.45, 2)
    xyz, box = _dev(engine, system, 0, 400)
    xyz, box = xyz.clone(), box.clone()
    xyz[100, 5] += 0.4                      # one atom far beyond the skin
    xyz[101, 700:720] -= 0.3                # a group of atoms
    xyz[250] = xyz[250].flip(0)             # every atom somewhere else
    box[333, 0, 0] += 0.125                 # another unit cell in one frame
    kicked = [100, 101, 250, 333]
    cells_a, cells_r, _ = _accumulate(engine, xyz, box, 0)
    list_a, list_r, st = _accumulate(engine, xyz, box, 2, audit=1)
    assert st["vl_valid"] == 1 and st["vl_tripped"] >= len(kicked)
    assert engine.get_stat("vl_disagree") == 0
    assert torch.equal(cells_a, list_a) and torch.equal(cells_r, list_r)
    # the kicked frames themselves against the oracle (they went through the fall back)
    xs, bs = xyz[kicked].contiguous(), box[kicked].contiguous()
    want_a, want_r = _oracle_tables(engine, system, role, xs, bs, len(kicked), 0.45, 2)
    rest = [f for f in range(400) if f not in kicked]
    rest_a, rest_r, _ = _accumulate(engine, xyz[rest].contiguous(), box[rest].contiguous(), 2)
    a, r = _tables_np(engine, list_a - rest_a, list_r - rest_r)
    assert np.array_equal(a, want_a) and np.array_equal(r, want_r)
    # NaN coordinates trip as well (and never hang / crash)
    xyz[7, 2000, 0] = float("nan")
    nan_a, nan_r, st = _accumulate(engine, xyz, box, 2)
    cells_a, cells_r, _ = _accumulate(engine, xyz, box, 0)
    assert st["vl_tripped"] >= len(kicked) + 1
    assert torch.equal(cells_a, nan_a) and torch.equal(cells_r, nan_r)


@pytest.mark.parametrize("why", ["small_box", "moving", "boxes"])
def test_unusable_list_falls_back_completely(engine, why):
    """A unit cell with fewer than three list cells per dimension, atoms that move further than
    any sensible skin, per-frame unit cells: the builder declares the list unusable (or every
    frame trips) and the result is still the oracle's."""
    rng = np.random.default_rng(41)
    if why == "small_box":
        system = synth.random_system(rng, 300, n_chains=2, box_kind="ortho", extent=1.5)
    else:
        system = synth.random_system(rng, 500, n_chains=2, box_kind="triclinic", extent=2.4)
    res, chain, role = system_arrays(system)
    engine.set_system(res, chain, role, 0.45, 2)
    frames = 24
    xyz, box = _dev(engine, system, 0, frames)
    if why == "moving":
        drift = torch.linspace(0.0, 3.0, frames, device=xyz.device).reshape(frames, 1, 1)
        xyz = (xyz + drift * torch.rand((1, xyz.shape[1], 3), device=xyz.device)).contiguous()
    if why == "boxes":
        scale = torch.linspace(1.0, 1.05, frames, device=xyz.device).reshape(frames, 1, 1)
        box = (box * scale).contiguous()
    got_a, got_r, st = _accumulate(engine, xyz, box, 2)
    assert st["vl_used"] == 1
    assert st["vl_valid"] == 0 or st["vl_tripped"] > 0
    want_a, want_r = _oracle_tables(engine, system, role, xyz, box, frames, 0.45, 2)
    a, r = _tables_np(engine, got_a, got_r)
    assert np.array_equal(a, want_a) and np.array_equal(r, want_r)


def test_public_api_uses_the_pair_list_for_long_trajectories(engine):
    """ContactFrequency of a 300-frame host trajectory goes through the pair-list path (batch >=
    256 frames) and equals the reference algorithm restated in the oracle."""
    import warnings
    warnings.simplefilter("ignore", PendingDeprecationWarning)
    rng = np.random.default_rng(43)
    system = synth.random_system(rng, 500, n_chains=2, box_kind="ortho", extent=2.3, h_fraction=0.0)
    traj = system.trajectory(300)
    query, hay = list(range(0, 300)), list(range(150, 500))
    freq = cm.ContactFrequency(traj, query=query, haystack=hay, cutoff=0.45, n_neighbors_ignored=2)
    assert engine.get_stat("vl_used") == 1 and engine.get_stat("vl_valid") == 1
    top = system.topology
    want_atoms, want_res = refpath.closed_form_frequency(traj.xyz, traj.unitcell_vectors, top.atom_residue,
                                                         top.residue_chain, query, hay, 0.45, 2)
    assert freq._atom_contacts == want_atoms and freq._residue_contacts == want_res


def test_mid_sized_system_many_tiles(engine):
    """7 000 atoms: several tiles, the guard pass decides per frame; against the oracle."""
    system = synth.solvated_triclinic(n_atoms=7000, lattice=20)
    res, chain, role = system_arrays(system)
    engine.set_system(res, chain, role, 0.45, 2)
    assert engine.path_kind == 1
    xyz, box = _dev(engine, system, 0, 40)
    xyz = xyz.clone()
    xyz[17, 3000:3010] += 0.5               # one frame trips
    got_a, got_r, st = _accumulate(engine, xyz, box, 2, audit=1)
    assert st["vl_valid"] == 1 and st["vl_tiles"] > 1 and st["vl_tripped"] == 1
    assert engine.get_stat("vl_disagree") == 0
    check = 6
    sel = [0, 1, 16, 17, 18, 39]
    xs, bs = xyz[sel].contiguous(), box[sel].contiguous()
    sub_a, sub_r, _ = _accumulate(engine, xs, bs, 2)
    want_a, want_r = _oracle_tables(engine, system, role, xs, bs, check, 0.45, 2)
    a, r = _tables_np(engine, sub_a, sub_r)
    assert np.array_equal(a, want_a) and np.array_equal(r, want_r)
    cells_a, cells_r, _ = _accumulate(engine, xyz, box, 0)
    assert torch.equal(cells_a, got_a) and torch.equal(cells_r, got_r)


# ---------------------------------------------------------------------------------------
# BASELINE.json config 4 / 5 at full system size (100 000 atoms)
# ---------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def s3(engine):
    system = synth.solvated_triclinic(n_atoms=100000)
    res, chain, role = system_arrays(system)
    return system, res, chain, role


def test_s3_full_size_frames_match_oracle(engine, s3):
    """S3 (100 000 atoms, triclinic): per-frame contact keys of frames {0, 1, F/2, F-1} of the
    1 000 000-frame trajectory against the oracle's cell-pruned exact enumeration, and the count
    tables of an 8-frame prefix through BOTH schedules (cell kernels, pair list) against the
    oracle's sparse sums."""
    system, res, chain, role = s3
    n = system.n_atoms
    F = 1000000
    engine.set_system(res, chain, role, 0.45, 2)
    assert engine.path_kind == 1
    rc = engine.n_res_c
    assert np.array_equal(engine.residue_map, np.arange(rc))          # every residue is used
    for f in (0, 1, F // 2, F - 1):
        xyz, box = _dev(engine, system, f, 1)
        akeys, rkeys = engine.frame_contacts(xyz, box, frame0=0)
        want = clib.contacts_frame(xyz[0].cpu().numpy(), system.box, 0.45, res, chain, role, 2, cells=True)
        frame, lo, hi = keys_to_frames(akeys)
        got = np.stack([lo, hi], axis=1)
        assert np.all(frame == 0)
        assert np.array_equal(got, want.astype(np.int64))
        _, rlo, rhi = keys_to_frames(rkeys)
        rl = np.minimum(res[want[:, 0]], res[want[:, 1]]).astype(np.int64)
        rh = np.maximum(res[want[:, 0]], res[want[:, 1]]).astype(np.int64)
        want_r = np.unique(rl * rc + rh)
        assert np.array_equal(rlo * rc + rhi, want_r)
    # 8-frame prefix: sparse tables
    xyz, box = _dev(engine, system, 0, 8)
    xh = xyz.cpu().numpy()
    all_flat, all_res = [], []
    for f in range(8):
        want = clib.contacts_frame(xh[f], system.box, 0.45, res, chain, role, 2, cells=True)
        all_flat.append(want[:, 0].astype(np.int64) * n + want[:, 1].astype(np.int64))
        rl = np.minimum(res[want[:, 0]], res[want[:, 1]]).astype(np.int64)
        rh = np.maximum(res[want[:, 0]], res[want[:, 1]]).astype(np.int64)
        all_res.append(np.unique(rl * rc + rh))
    want_idx, want_cnt = np.unique(np.concatenate(all_flat), return_counts=True)
    want_ridx, want_rcnt = np.unique(np.concatenate(all_res), return_counts=True)
    atom_t, res_t = engine.new_tables()
    try:
        for mode in (0, 2):
            atom_t.zero_()
            res_t.zero_()
            engine.set_option("verlet", mode)
            engine.accumulate(xyz, box, atom_t, res_t)
            torch.cuda.synchronize()
            if mode == 2:
                assert engine.get_stat("vl_used") == 1 and engine.get_stat("vl_valid") == 1
                assert engine.get_stat("vl_tiles") > 100 and engine.get_stat("vl_tripped") == 0
            (idx, cnt), (ridx, rcnt) = engine.export_many([atom_t, res_t])
            assert np.array_equal(idx, want_idx) and np.array_equal(cnt, want_cnt.astype(np.uint32))
            assert np.array_equal(ridx, want_ridx) and np.array_equal(rcnt, want_rcnt.astype(np.uint32))
    finally:
        engine.set_option("verlet", 1)
        del atom_t, res_t
        torch.cuda.empty_cache()


def test_s4_full_size_min_dist_and_nearest_match_oracle(engine, s3):
    """S4 on the 100 000-atom system: MinimumDistanceCounter (30-atom ligand x all solvent, the
    product; and an arbitrary pair list through _compute_from_atom_pairs) bit-equal to the oracle's
    compute_distances restatement; NearestAtoms for 400 sampled atoms against the oracle's
    neighbour list + distances."""
    system, res, chain, role = s3
    n = system.n_atoms
    frames = 3
    xyz, box = _dev(engine, system, 0, frames)
    xh, bh = xyz.cpu().numpy(), box.cpu().numpy()
    q = np.arange(0, 30)
    h = np.arange(2700, n)
    dmin, darg = engine.min_dist(xyz, box, q, h, False)
    want_d, want_i = clib.min_dist(xh, q, h, bh, False)
    assert np.array_equal(dmin.cpu().numpy(), want_d) and np.array_equal(darg.cpu().numpy(), want_i)
    # arbitrary pair list (the reference's seam, min_dist.py:142-168)
    rng = np.random.default_rng(9)
    pairs = np.stack([rng.integers(0, n, 20000), rng.integers(0, n, 20000)], axis=1)
    pairs = pairs[pairs[:, 0] != pairs[:, 1]]
    traj = cm.Trajectory(xh, system.topology, bh)
    got_d, got_i = cm.MinimumDistanceCounter._compute_from_atom_pairs(traj, pairs)
    dist = clib.distances(xh, pairs, bh, False)
    assert np.array_equal(got_d, dist.min(axis=1)) and np.array_equal(got_i, dist.argmin(axis=1))
    # NearestAtoms: all 100 000 atoms on the device, 400 of them checked against the oracle
    near = cm.NearestAtoms(traj, cutoff=0.45, frame_number=0)
    neighbours = clib.neighborlist(xh[0], bh[0], 0.45, cells=True)
    sample = rng.choice(n, size=400, replace=False)
    for atom in sample:
        allowed = [int(j) for j in neighbours[atom] if res[j] != res[atom]]
        if not allowed:
            assert int(atom) not in near.nearest
            continue
        d = clib.distances(xh[:1], np.array([(atom, j) for j in allowed], dtype=np.int32), bh[:1], False)[0]
        best = sorted(zip(d.tolist(), allowed))[0]
        assert near.nearest[int(atom)] == best[1]
        assert near.nearest_distance[int(atom)] == np.float32(best[0])


def test_streamed_host_schedule_head_and_tail_chunks(engine):
    """A pinned host trajectory long enough for the automatic chunk schedule of
    cmb_accumulate_host (short first chunk that the list is built from, quarter-size tail chunk cut
    into fine work items, unit cells copied once): same tables as the device entry point, with and
    without frames that trip the guard inside the head and the tail chunk."""
    rng = np.random.default_rng(77)
    system = synth.random_system(rng, 300, n_chains=2, box_kind="ortho", extent=1.9)
    res, chain, role = system_arrays(system)
    engine.set_system(res, chain, role, 0.45, 2)
    assert engine.path_kind == 0
    F = 8000
    xyz, box = _dev(engine, system, 0, F)
    xyz, box = xyz.clone(), box.clone()
    fill = engine.get_stat("sm_count") * max(1, engine.get_stat("ctas_per_sm")) * 2
    chunk = max(fill, (F + 3) // 4)                       # 80 MB would hold far more frames of 300 atoms
    head = max(fill, chunk // 4)
    want = ([head] if F >= 3 * chunk and chunk >= 2 * head else [])
    rest = F - sum(want)
    want += [chunk] * (rest // chunk) + ([rest % chunk] if rest % chunk else [])
    tail = max(fill, want[-1] // 4)
    if want[-1] >= 2 * tail:
        want[-1:] = [want[-1] - tail, tail]
    for kicked in ((), (3, F - 5, F // 2)):
        for f in kicked:
            xyz[f, 11] += 0.5
        dev_a, dev_r, _ = _accumulate(engine, xyz, box, 1)
        h_xyz = torch.empty(xyz.shape, dtype=torch.float32, pin_memory=True).copy_(xyz)
        h_box = torch.empty(box.shape, dtype=torch.float32, pin_memory=True).copy_(box)
        host_a, host_r = engine.new_tables()
        engine.set_option("stage_trace", 1)
        try:
            engine.accumulate_host(h_xyz.numpy(), h_box.numpy(), host_a, host_r)
            assert engine.get_stat("trace_chunks") == len(want)
            done = [engine.get_stat("trace_done_%d" % c) for c in range(len(want))]
            assert all(d > 0 for d in done)
        finally:
            engine.set_option("stage_trace", 0)
        assert engine.get_stat("vl_used") == 1
        assert torch.equal(host_a, dev_a) and torch.equal(host_r, dev_r)
    assert len(want) >= 5 and want[0] < want[1] and want[-1] < want[-2] + want[-1]


def test_builder_graph_equals_plain_launches(engine):
    """The list builder replayed as one CUDA graph (default) against the same launches issued one
    by one: identical tables, on a single-tile and on a many-tile system, across repeated calls
    with different batches (the per-call arguments are read from device memory) and after a change
    of system (the graph is captured again)."""
    for system, frames in ((synth.kras_like(), 300), (synth.solvated_triclinic(n_atoms=7000, lattice=20), 40)):
        res, chain, role = system_arrays(system)
        engine.set_system(res, chain, role, 0.45, 2)
        for f0 in (0, 1000):
            xyz, box = _dev(engine, system, f0, frames)
            engine.set_option("verlet_graph", 0)
            try:
                plain_a, plain_r, st0 = _accumulate(engine, xyz, box, 2)
            finally:
                engine.set_option("verlet_graph", 1)
            graph_a, graph_r, st1 = _accumulate(engine, xyz, box, 2)
            assert st0["vl_valid"] == 1 and st1["vl_valid"] == 1 and st0["vl_entries"] == st1["vl_entries"]
            assert engine.get_stat("vl_graph_state") == 1
            assert torch.equal(plain_a, graph_a) and torch.equal(plain_r, graph_r)


def test_optimistic_tiles_overflow_rebuilds_the_list(engine):
    """The tile edge comes from the occupancy of the list cells, counting only a share of every
    tile's shell atoms as slots; with that share set far too low the tiles overflow and the list is
    built again with the whole shell counted -- synchronously for a device batch, one chunk later
    (never waiting for the device) for a streamed host trajectory, whose first frames meanwhile run
    through the fall back.  Same tables as the cell kernels either way.  Atoms that fill one corner of
    a larger unit cell (2.2 x the volume) get a usable list as well: the estimate looks at the
    fullest tile, not at the average density."""
    system = synth.solvated_triclinic(n_atoms=20000, lattice=28)
    system.image_shifts = False                           # (the shifts are lattice vectors of the ORIGINAL cell)
    res, chain, role = system_arrays(system)
    engine.set_system(res, chain, role, 0.45, 2)
    assert engine.path_kind == 1
    F = 96
    xyz, box = _dev(engine, system, 0, F)
    box = (box * 1.3).contiguous()
    cells_a, cells_r, _ = _accumulate(engine, xyz, box, 0)
    try:
        engine.set_option("verlet_conservative_tiles", 0)
        before = engine.get_stat("vl_rebuilds")
        list_a, list_r, st = _accumulate(engine, xyz, box, 2)
        assert st["vl_valid"] == 1 and st["vl_tripped"] == 0 and st["vl_tiles"] > 1
        assert engine.get_stat("vl_rebuilds") == before and engine.get_stat("vl_conservative") == 0
        assert torch.equal(cells_a, list_a) and torch.equal(cells_r, list_r)
        tiles_optimistic = st["vl_tiles"]
        # an estimate that is far too optimistic: overflow, pessimistic rebuild
        engine.set_option("verlet_tile_share_pct", 0)
        list_a, list_r, st = _accumulate(engine, xyz, box, 2)
        assert engine.get_stat("vl_rebuilds") == before + 1 and engine.get_stat("vl_conservative") == 1
        assert engine.get_stat("vl_rebuild_code") in (9, 10, 11, 12, 13)
        assert st["vl_valid"] == 1 and st["vl_tripped"] == 0 and st["vl_tiles"] >= tiles_optimistic
        assert torch.equal(cells_a, list_a) and torch.equal(cells_r, list_r)
        # streamed: chunks of 24 frames share one list; the first ones meet the overflowing list
        engine.set_option("verlet_conservative_tiles", 0)
        before = engine.get_stat("vl_rebuilds")
        h_xyz = torch.empty(xyz.shape, dtype=torch.float32, pin_memory=True).copy_(xyz)
        h_box = torch.empty(box.shape, dtype=torch.float32, pin_memory=True).copy_(box)
        host_a, host_r = engine.new_tables()
        engine.set_option("verlet", 2)
        engine.accumulate_host(h_xyz.numpy(), h_box.numpy(), host_a, host_r, chunk_frames=24)
        torch.cuda.synchronize()
        assert engine.get_stat("vl_rebuilds") == before + 1 and engine.get_stat("vl_conservative") == 1
        assert engine.get_stat("vl_valid") == 1
        assert torch.equal(cells_a, host_a) and torch.equal(cells_r, host_r)
    finally:
        engine.set_option("verlet", 1)
        engine.set_option("verlet_tile_share_pct", 60)
        engine.set_option("verlet_conservative_tiles", 0)
