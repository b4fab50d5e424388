"""GPU: the global-memory cell path, the device generator, the boundary report and
size-independent properties at BASELINE.json's full sizes."""
import numpy as np
import pytest
import torch

from contact_map_b200 import synth
from oracle import clib
from tests.helpers import system_arrays, dense_from_export, keys_to_frames

pytestmark = pytest.mark.gpu


def _dev(engine, system, frame0, n_frames):
    dev = engine.device
    xyz = engine.synth_frames(torch.as_tensor(system.base, device=dev),
                              torch.as_tensor(system.groups.astype(np.int32), device=dev),
                              None if system.box is None else torch.as_tensor(system.box.reshape(9), device=dev),
                              system.seed, system.amplitude, system.image_shifts, frame0, n_frames)
    box = None if system.box is None else \
        torch.as_tensor(system.box, device=dev).reshape(1, 3, 3).expand(n_frames, 3, 3).contiguous()
    return xyz, box


@pytest.mark.parametrize("which", ["s1", "s3", "open"])
def test_device_generator_bit_identical_to_host(engine, which):
    if which == "s1":
        system = synth.kras_like()
    elif which == "s3":
        system = synth.solvated_triclinic(n_atoms=5000, lattice=18)
    else:
        system = synth.random_system(np.random.default_rng(2), 300, box_kind="none")
    xyz, _ = _dev(engine, system, 37, 5)
    host = system.frames(37, 5)
    assert np.array_equal(xyz.cpu().numpy().view(np.uint32), host.view(np.uint32))


@pytest.mark.parametrize("box_kind", ["ortho", "triclinic", "none"])
def test_large_path_forced_on_small_systems(engine, box_kind):
    """The global-memory cell path (kl_* kernels) against the oracle and the fused kernel."""
    rng = np.random.default_rng(77)
    system = synth.random_system(rng, 700, n_chains=3, box_kind=box_kind, extent=2.4)
    query = np.sort(rng.choice(700, size=500, replace=False))
    hay = np.sort(rng.choice(700, size=600, replace=False))
    res, chain, role = system_arrays(system, query, hay)
    n_frames = 4
    xyz = system.frames(0, n_frames)
    box = None if system.box is None else np.broadcast_to(system.box, (n_frames, 3, 3)).copy()
    ref_atom, ref_res = clib.accumulate_dense(xyz, box, 0.45, res, chain, role, 2, system.topology.n_residues)
    d_xyz = torch.as_tensor(xyz, device=engine.device)
    d_box = None if box is None else torch.as_tensor(box, device=engine.device)
    results = []
    # global-memory path one frame at a time (frame stamps), batched (per-frame bitmaps), fused kernel
    for force, batch in ((1, 0), (1, 2), (0, 1)):
        engine.set_option("force_path", force)
        engine.set_option("large_batch", batch)
        engine.set_system(res, chain, role, 0.45, 2)
        assert engine.path_kind == force
        atom_t, res_t = engine.new_tables()
        engine.accumulate(d_xyz, d_box, atom_t, res_t)
        idx, cnt = engine.export_nonzero(atom_t)
        got = dense_from_export(idx, cnt, 700)
        assert np.array_equal(got, ref_atom)
        ridx, rcnt = engine.export_nonzero(res_t)
        rmap = engine.residue_map
        assert np.array_equal(dense_from_export(ridx, rcnt, engine.n_res_c), ref_res[np.ix_(rmap, rmap)])
        akeys, rkeys = engine.frame_contacts(d_xyz, d_box)
        results.append((akeys, rkeys))
        assert len(akeys) == int(ref_atom.sum()) and len(rkeys) == int(ref_res.sum())
    engine.set_option("force_path", -1)
    engine.set_option("large_batch", 1)
    for other in results[1:]:
        assert np.array_equal(results[0][0], other[0]) and np.array_equal(results[0][1], other[1])


@pytest.mark.parametrize("box_kind", ["triclinic", "none"])
def test_hashed_tables_match_dense(engine, box_kind):
    """Hashed count tables (cmb_set_table_mode) of the global-memory cell path: same counts as
    the oracle's dense tables, through growth (tiny initial tables are rehashed several times),
    through the host entry point, and mixed with a dense residue table."""
    from contact_map_b200.engine import HashTable
    rng = np.random.default_rng(78)
    system = synth.random_system(rng, 900, n_chains=2, box_kind=box_kind, extent=2.6)
    res, chain, role = system_arrays(system)
    n_frames = 7
    xyz = system.frames(0, n_frames)
    box = None if system.box is None else np.broadcast_to(system.box, (n_frames, 3, 3)).copy()
    ref_atom, ref_res = clib.accumulate_dense(xyz, box, 0.45, res, chain, role, 2, system.topology.n_residues)
    engine.set_option("force_path", 1)
    try:
        engine.set_system(res, chain, role, 0.45, 2)
        rmap = engine.residue_map
        want_res = ref_res[np.ix_(rmap, rmap)]
        d_xyz = torch.as_tensor(xyz, device=engine.device)
        d_box = None if box is None else torch.as_tensor(box, device=engine.device)
        for res_hashed, host in ((True, False), (False, False), (True, True)):
            atom_t = HashTable(5, engine.device)                 # 32 slots: forces repeated growth
            res_t = HashTable(5, engine.device) if res_hashed else engine.new_tables(hashed=False)[1]
            if host:
                engine.accumulate_host(xyz, box, atom_t, res_t, chunk_frames=2)
            else:
                engine.accumulate(d_xyz, d_box, atom_t, res_t)
            assert atom_t.log2 > 5 and atom_t.used == np.count_nonzero(ref_atom)
            assert 100 * atom_t.used <= 85 * atom_t.numel()
            (idx, cnt), (ridx, rcnt) = engine.export_many([atom_t, res_t])
            assert np.all(np.diff(idx) > 0)
            assert np.array_equal(dense_from_export(idx, cnt, 900), ref_atom)
            assert np.array_equal(dense_from_export(ridx, rcnt, engine.n_res_c), want_res)
        # the default for this size stays dense; hashed=True gives sensibly sized tables
        a, r = engine.new_tables()
        assert isinstance(a, torch.Tensor) and isinstance(r, torch.Tensor)
        a, r = engine.new_tables(hashed=True)
        assert isinstance(a, HashTable) and a.numel() >= 32 * 900
    finally:
        engine.set_option("force_path", -1)
    # the fused path refuses hashed tables loudly
    engine.set_system(res, chain, role, 0.45, 2)
    assert engine.path_kind == 0
    with pytest.raises(Exception):
        engine.accumulate(d_xyz, d_box, HashTable(10, engine.device), engine.new_tables()[1])
    engine._set_table_mode(None, None)


def test_large_system_matches_oracle(engine):
    """n_used > 6144: S3-like solvated triclinic system (12 000 atoms), two frames."""
    system = synth.solvated_triclinic(n_atoms=12000, lattice=23)
    res, chain, role = system_arrays(system)
    n = system.n_atoms
    engine.set_system(res, chain, role, 0.45, 2)
    assert engine.path_kind == 1
    xyz, box = _dev(engine, system, 0, 2)
    akeys, rkeys = engine.frame_contacts(xyz, box)
    frame, lo, hi = keys_to_frames(akeys)
    h_xyz = xyz.cpu().numpy()
    for f in range(2):
        want = clib.contacts_frame(h_xyz[f], system.box, 0.45, res, chain, role, 2, cells=True)
        sel = frame == f
        assert np.array_equal(np.stack([lo[sel], hi[sel]], axis=1), want)
        assert len(want) > 50000
    # tables == sum of the per-frame keys, dense and hashed
    flat, counts = np.unique(lo * n + hi, return_counts=True)
    for hashed in (False, True):
        atom_t, res_t = engine.new_tables(hashed=hashed)
        engine.accumulate(xyz, box, atom_t, res_t)
        (idx, cnt), (ridx, rcnt) = engine.export_many([atom_t, res_t])
        assert int(cnt.sum()) == len(akeys) and int(rcnt.sum()) == len(rkeys)
        assert np.array_equal(idx, flat) and np.array_equal(cnt, counts.astype(np.uint32))


def test_boundary_report_matches_oracle(engine):
    """north_star: pairs within 1 ulp of the cutoff are enumerated.  A dense jittered system has
    a few such pairs at ulps=64; the lists must agree exactly with the oracle's."""
    system = synth.kras_like()
    res, chain, role = system_arrays(system)
    engine.set_system(res, chain, role, 0.45, 2)
    n_frames = 6
    xyz, box = _dev(engine, system, 0, n_frames)
    h = xyz.cpu().numpy()
    for ulps in (1, 64):
        rec = engine.boundary_pairs(xyz, box, ulps=ulps)
        want = []
        for f in range(n_frames):
            a, b, d2 = clib.boundary_pairs(h[f], system.box, 0.45, res, chain, role, 2, ulps=ulps)
            want.extend((f, int(x), int(y), float(z)) for x, y, z in zip(a, b, d2))
        got = [(int(r["frame"]), int(r["a"]), int(r["b"]), float(r["d2"])) for r in rec]
        assert got == sorted(want)
    assert len(got) > 0


def test_full_size_properties_s1(engine):
    """BASELINE configs[1] at full size (10 000 frames): partition invariance, host == device,
    table mass == number of per-frame keys, oracle spot checks."""
    system = synth.kras_like()
    res, chain, role = system_arrays(system)
    n = system.n_atoms
    engine.set_system(res, chain, role, 0.45, 2)
    F = 10000
    xyz, box = _dev(engine, system, 0, F)
    total_a, total_r = engine.new_tables()
    engine.accumulate(xyz, box, total_a, total_r)
    # (1) partition invariance: three uneven shards into one pair of tables
    part_a, part_r = engine.new_tables()
    for a, b in ((0, 1), (1, 4097), (4097, F)):
        engine.accumulate(xyz[a:b], box[a:b], part_a, part_r)
    assert torch.equal(part_a, total_a) and torch.equal(part_r, total_r)
    # (2) frame order does not matter
    perm = torch.randperm(F, device=engine.device, generator=torch.Generator(engine.device).manual_seed(1))
    perm_a, perm_r = engine.new_tables()
    engine.accumulate(xyz[perm].contiguous(), box[perm].contiguous(), perm_a, perm_r)
    assert torch.equal(perm_a, total_a) and torch.equal(perm_r, total_r)
    # (3) host-buffer entry point == device entry point
    host_a, host_r = engine.new_tables()
    engine.accumulate_host(xyz.cpu().numpy(), box.cpu().numpy(), host_a, host_r)
    assert torch.equal(host_a, total_a) and torch.equal(host_r, total_r)
    # (4) mass: sum of the tables == number of per-frame keys; counts bounded by F
    akeys, rkeys = engine.frame_contacts(xyz[:2000], box[:2000])
    sub_a, sub_r = engine.new_tables()
    engine.accumulate(xyz[:2000], box[:2000], sub_a, sub_r)
    assert int(sub_a.sum().item()) == len(akeys) and int(sub_r.sum().item()) == len(rkeys)
    assert int(total_a.max().item()) <= F and int(total_r.max().item()) <= F
    # (5) strictly upper-triangular, nothing inside the exclusion band
    idx, cnt = engine.export_nonzero(total_a)
    lo, hi = idx // n, idx % n
    assert np.all(lo < hi) and np.all(np.abs(res[lo] - res[hi]) > 2)
    # (6) oracle on sampled frames
    h = xyz[[0, 1, F // 2, F - 1]].cpu().numpy()
    frame, klo, khi = keys_to_frames(engine.frame_contacts(xyz[[0, 1, F // 2, F - 1]].contiguous(),
                                                           box[:4].contiguous())[0])
    for f in range(4):
        want = clib.contacts_frame(h[f], system.box, 0.45, res, chain, role, 2, cells=True)
        sel = frame == f
        assert np.array_equal(np.stack([klo[sel], khi[sel]], axis=1), want)
    # (7) both count tables of a 96-frame block (frames 5000..5095) against the oracle's dense tables
    blk = slice(5000, 5096)
    ref_atom, ref_res = clib.accumulate_dense(xyz[blk].cpu().numpy(), box[blk].cpu().numpy(), 0.45, res, chain,
                                              role, 2, system.topology.n_residues)
    blk_a, blk_r = engine.new_tables()
    engine.accumulate(xyz[blk], box[blk], blk_a, blk_r)
    (idx, cnt), (ridx, rcnt) = engine.export_many([blk_a, blk_r])
    assert np.array_equal(dense_from_export(idx, cnt, n), ref_atom)
    rmap = engine.residue_map
    assert np.array_equal(dense_from_export(ridx, rcnt, engine.n_res_c), ref_res[np.ix_(rmap, rmap)])


def test_binding_system_s2_matches_oracle(engine):
    """BASELINE configs[2]: query = partner chain, haystack = receptor (first 48 frames vs oracle)."""
    system = synth.binding_system()
    res, chain, role = system_arrays(system, system.query, system.haystack)
    engine.set_system(res, chain, role, 0.45, 2)
    F = 48
    xyz, box = _dev(engine, system, 0, F)
    atom_t, res_t = engine.new_tables()
    engine.accumulate(xyz, box, atom_t, res_t)
    idx, cnt = engine.export_nonzero(atom_t)
    n = system.n_atoms
    h_xyz, h_box = xyz.cpu().numpy(), box.cpu().numpy()
    ref_atom, ref_res = clib.accumulate_dense(h_xyz, h_box, 0.45, res, chain, role, 2,
                                              system.topology.n_residues)
    assert np.array_equal(dense_from_export(idx, cnt, n), ref_atom)
    assert ref_atom.sum() > 100
    # every contact joins the partner (query) with the receptor (haystack)
    lo, hi = idx // n, idx % n
    assert np.all((role[lo] | role[hi]) == 3) and np.all(role[lo] != role[hi])


@pytest.mark.parametrize("ligand", [True, False])
def test_s2_contact_trajectory_per_frame_keys_match_oracle(engine, ligand):
    """BASELINE configs[2] at benchmark SHAPE (30-atom ligand / 600-atom partner chain against the
    2 700-atom receptor) through the public API from HOST memory: the per-frame atom and residue
    sets of ContactTrajectory equal the oracle's for every frame of a 40-frame block, and
    ContactDifference of two halves equals the difference of the oracle's count tables."""
    import contact_map_b200 as cm
    system = synth.binding_system(ligand=ligand)
    top = system.topology
    res, chain, role = system_arrays(system, system.query, system.haystack)
    F = 40
    xyz = system.frames(0, F)
    box = np.broadcast_to(system.box, (F, 3, 3)).copy()
    traj = cm.Trajectory(xyz, top, box)
    kw = dict(query=system.query, haystack=system.haystack, cutoff=0.45, n_neighbors_ignored=2)
    ct = cm.ContactTrajectory(traj, **kw)
    atoms_csr, res_csr = ct._csr
    n_hits = 0
    for f in range(F):
        pairs = clib.contacts_frame(xyz[f], box[f], 0.45, res, chain, role, 2)
        want = {(int(a), int(b)) for a, b in pairs}
        a, b = atoms_csr.ptr[f], atoms_csr.ptr[f + 1]
        got = set(zip(atoms_csr.i[a:b].tolist(), atoms_csr.j[a:b].tolist()))
        assert got == want, f
        want_res = {(min(int(res[a_]), int(res[b_])), max(int(res[a_]), int(res[b_]))) for a_, b_ in pairs}
        a, b = res_csr.ptr[f], res_csr.ptr[f + 1]
        assert set(zip(res_csr.i[a:b].tolist(), res_csr.j[a:b].tolist())) == want_res, f
        n_hits += len(want)
        # every contact joins the query with the haystack
        assert all((role[i] | role[j]) == 3 and role[i] != role[j] for i, j in want)
    assert n_hits > (200 if ligand else 100)
    # a frame object built from the CSR arrays holds the same set
    assert set(map(tuple, map(sorted, ct[3]._atom_contacts.keys()))) == \
        {(int(a), int(b)) for a, b in clib.contacts_frame(xyz[3], box[3], 0.45, res, chain, role, 2)}
    # ContactDifference(first half, second half) == difference of the oracle tables
    half = F // 2
    diff = cm.ContactFrequency(traj[:half], **kw) - cm.ContactFrequency(traj[half:], **kw)
    ref_a0, _ = clib.accumulate_dense(xyz[:half], box[:half], 0.45, res, chain, role, 2, top.n_residues)
    ref_a1, _ = clib.accumulate_dense(xyz[half:], box[half:], 0.45, res, chain, role, 2, top.n_residues)
    want_diff = ref_a0.astype(np.float64) / half - ref_a1.astype(np.float64) / half
    got = diff.atom_contacts.counter
    ii, jj = np.nonzero(np.triu(ref_a0 + ref_a1))
    assert len(got) == len(ii)
    for i, j in zip(ii.tolist(), jj.tolist()):
        assert got[frozenset((i, j))] == want_diff[i, j]


def test_s2_long_trajectory_properties(engine):
    """S2 ligand at its full length (100 000 frames from host memory, the configuration north_star
    names): size-independent properties -- the per-frame keys summed over frames equal the
    ContactFrequency tables of the same trajectory, frame pointers are monotone and complete, and a
    sampled set of frames equals the oracle."""
    import contact_map_b200 as cm
    system = synth.binding_system(ligand=True)
    top = system.topology
    res, chain, role = system_arrays(system, system.query, system.haystack)
    F = 100000
    d_xyz, d_box = _dev(engine, system, 0, F)
    xyz, box = d_xyz.cpu().numpy(), d_box.cpu().numpy()
    del d_xyz, d_box
    traj = cm.Trajectory(xyz, top, box)
    kw = dict(query=system.query, haystack=system.haystack, cutoff=0.45, n_neighbors_ignored=2)
    ct = cm.ContactTrajectory(traj, **kw)
    assert len(ct) == F
    atoms_csr, res_csr = ct._csr
    for csr in (atoms_csr, res_csr):
        assert csr.ptr[0] == 0 and csr.ptr[-1] == csr.i.shape[0] and np.all(np.diff(csr.ptr) >= 0)
        assert np.all(csr.i < csr.j)
    freq = cm.ContactFrequency(traj, **kw)
    total = ct.contact_frequency()
    assert total._atom_contacts == freq._atom_contacts
    assert total._residue_contacts == freq._residue_contacts
    for f in (0, 1, F // 2, F - 1):
        pairs = clib.contacts_frame(xyz[f], box[f], 0.45, res, chain, role, 2)
        a, b = atoms_csr.ptr[f], atoms_csr.ptr[f + 1]
        assert set(zip(atoms_csr.i[a:b].tolist(), atoms_csr.j[a:b].tolist())) == {(int(p), int(q)) for p, q in pairs}


@pytest.mark.parametrize("box_kind,force_path,max_shift", [("ortho", 0, 2), ("triclinic", 0, 2), ("none", 0, 0),
                                                           ("triclinic", 1, 2), ("ortho", 0, 400)])
def test_pairs_at_the_cutoff_are_adjudicated_exactly(engine, box_kind, force_path, max_shift):
    """The screening thresholds (frame_margins) must never decide a pair the reference arithmetic
    would decide differently: 1 500 atom pairs are placed within a few float steps of the cutoff
    (both sides), then moved by whole box vectors so that the raw coordinates are large and the
    rounding of the reference chain matters.  Every pair is its own two single-atom residues in
    two chains, so the oracle's contact set is exactly the set of pairs it judges inside.
    max_shift = 400 puts the atoms thousands of nm from the origin: the error bound exceeds the
    useful range and the kernel must fall back to adjudicating everything exactly."""
    rng = np.random.default_rng(2024)
    n_pairs, cutoff = 1500, 0.45
    side = 12
    if box_kind == "ortho":
        box = np.diag([14.4, 15.0, 13.8]).astype(np.float32)
    elif box_kind == "triclinic":
        box = np.array([[14.4, 0, 0], [3.1, 15.0, 0], [-2.7, 4.2, 13.8]], dtype=np.float32)
    else:
        box = None
    # pair centres on a coarse lattice (1.1 nm apart: different pairs never interact)
    sites = np.stack(np.meshgrid(*[np.arange(side)] * 3, indexing="ij"), axis=-1).reshape(-1, 3)[:n_pairs]
    centre = (sites * 1.1 + 0.6).astype(np.float64)
    direction = rng.normal(size=(n_pairs, 3))
    direction /= np.linalg.norm(direction, axis=1, keepdims=True)
    c32 = np.float32(cutoff)
    steps = rng.integers(-6, 7, size=n_pairs)                        # float steps of the DISTANCE
    dist = np.float64(c32) * (1.0 + steps * 2.0 ** -24)
    a = centre - 0.5 * dist[:, None] * direction
    b = centre + 0.5 * dist[:, None] * direction
    xyz = np.empty((3, 2 * n_pairs, 3), dtype=np.float32)
    for f in range(3):
        pa, pb = a.copy(), b.copy()
        if box is not None:
            # molecules "not made whole": each atom moved by its own lattice vector
            for p in (pa, pb):
                shift = rng.integers(-max_shift, max_shift + 1, size=(n_pairs, 3)).astype(np.float64)
                p += shift @ box.astype(np.float64)
        xyz[f, 0::2] = pa.astype(np.float32)
        xyz[f, 1::2] = pb.astype(np.float32)
    n = 2 * n_pairs
    res = np.arange(n, dtype=np.int32)
    chain = (np.arange(n) % 2).astype(np.int32)
    role = np.full(n, 3, dtype=np.uint8)
    boxes = None if box is None else np.broadcast_to(box, (3, 3, 3)).copy()
    ref_atom, _ = clib.accumulate_dense(xyz, boxes, cutoff, res, chain, role, 0, n)
    inside = int(ref_atom.sum())
    assert 0.1 * 3 * n_pairs < inside < 0.9 * 3 * n_pairs            # both outcomes well represented
    engine.set_option("force_path", force_path if force_path else -1)
    try:
        engine.set_system(res, chain, role, cutoff, 0)
        assert engine.path_kind == force_path
        atom_t, res_t = engine.new_tables(hashed=False)
        d_xyz = torch.as_tensor(xyz, device=engine.device)
        d_box = None if boxes is None else torch.as_tensor(boxes, device=engine.device)
        engine.accumulate(d_xyz, d_box, atom_t, res_t)
        idx, cnt = engine.export_nonzero(atom_t)
        assert np.array_equal(dense_from_export(idx, cnt, n), ref_atom)
        # and the boundary report sees them (north_star: pairs within N ulps are enumerated)
        rec = engine.boundary_pairs(d_xyz[:1], None if d_box is None else d_box[:1], ulps=64)
        wa, wb, wd = clib.boundary_pairs(xyz[0], box, cutoff, res, chain, role, 0, ulps=64)
        assert len(wa) > (n_pairs // 2 if max_shift <= 2 else 0)
        order = np.lexsort((wb, wa))
        assert np.array_equal(rec["a"], wa[order]) and np.array_equal(rec["b"], wb[order])
        assert np.array_equal(rec["d2"], wd[order])
    finally:
        engine.set_option("force_path", -1)


@pytest.mark.parametrize("case", ["s1", "triclinic_large_path", "varying_cells", "far_from_origin", "open", "at_the_cutoff"])
def test_screen_audit_finds_no_disagreement(engine, case):
    """the screening margins (frame_margins) are AUDITED, not only argued --
    the audit variant of the cell kernels runs the reference arithmetic on every stencil candidate
    that is not residue-excluded and counts the pairs on which the production thresholds alone
    (d2 <= c2_lo: contact, d2 >= c2_hi: no contact) would have decided differently: zero, on S1
    (> 10^9 pairs), the triclinic global-memory cell path, cells that change every frame,
    coordinates hundreds of nm from the origin (unwrapped trajectories, up to the 900 nm domain
    limit), no unit cell, and pairs placed within a few float steps of the cutoff."""
    rng = np.random.default_rng(77)
    dev = engine.device
    min_pairs = 1
    try:
        if case == "s1":
            system = synth.kras_like()
            res, chain, role = system_arrays(system)
            engine.set_system(res, chain, role, 0.45, 2)
            xyz, box = _dev(engine, system, 0, 8000)
            min_pairs = 10 ** 9
        elif case == "triclinic_large_path":
            system = synth.solvated_triclinic(n_atoms=12000, lattice=23)
            res, chain, role = system_arrays(system)
            engine.set_option("force_path", 1)
            engine.set_system(res, chain, role, 0.45, 2)
            xyz, box = _dev(engine, system, 0, 48)
            min_pairs = 10 ** 7
        else:
            kind = {"varying_cells": "triclinic", "far_from_origin": "ortho", "open": "open", "at_the_cutoff": "triclinic"}[case]
            system = synth.random_system(rng, 2500, n_chains=3, box_kind=kind, extent=3.2)
            res, chain, role = system_arrays(system)
            engine.set_system(res, chain, role, 0.45, 1)
            F = 200
            h_xyz = system.frames(0, F)
            h_box = None if system.box is None else np.broadcast_to(system.box, (F, 3, 3)).copy()
            if case == "varying_cells":
                h_box = (h_box * (1.0 + 0.002 * np.arange(F, dtype=np.float32))[:, None, None]).astype(np.float32)
            if case == "far_from_origin":
                # whole-cell shifts of up to +-200 cells per residue group: |coordinates| up to ~850 nm
                shift = rng.integers(-200, 201, size=(F, 1, 3)).astype(np.float32) * np.diag(system.box)[None, None, :]
                h_xyz = (h_xyz + shift).astype(np.float32)
                assert np.abs(h_xyz).max() > 500 and np.abs(h_xyz).max() < 900
            if case == "at_the_cutoff":
                # move every odd atom onto a sphere of radius cutoff (+- a few float steps) around its predecessor
                d = rng.normal(size=(F, 1250, 3))
                d /= np.linalg.norm(d, axis=2, keepdims=True)
                r = np.float32(0.45) * (1.0 + rng.integers(-6, 7, size=(F, 1250, 1)) * 6e-8)
                h_xyz[:, 1::2] = (h_xyz[:, 0::2].astype(np.float64) + d * r).astype(np.float32)
            xyz = torch.as_tensor(h_xyz, device=dev)
            box = None if h_box is None else torch.as_tensor(h_box, device=dev)
            min_pairs = 10 ** 6
        audited, bad = engine.screen_audit(xyz, box)
    finally:
        engine.set_option("force_path", -1)
    assert bad == 0
    assert audited >= min_pairs
