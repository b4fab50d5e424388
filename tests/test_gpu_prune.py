"""GPU: the cell-pruned NearestAtoms / MinimumDistanceCounter kernels (cmb_prune.cuh) against the
brute-force kernels and the oracle: identical distances (float32 bits) and identical arg-min / tie
decisions, on orthorhombic, triclinic and open systems, per-frame-varying cells, query and
haystack far apart (every frame re-done by brute force) and a mixture of both."""
import numpy as np
import pytest
import torch

from contact_map_b200 import synth
from oracle import clib

pytestmark = pytest.mark.gpu


def _system(kind, n_atoms, extent, seed):
    rng = np.random.default_rng(seed)
    system = synth.random_system(rng, n_atoms, n_chains=3, box_kind=kind, extent=extent)
    return system, rng


def _frames(engine, system, n_frames):
    xyz = system.frames(0, n_frames)
    box = None if system.box is None else np.broadcast_to(system.box, (n_frames, 3, 3)).copy()
    dev = engine.device
    return (xyz, box, torch.as_tensor(xyz, device=dev),
            None if box is None else torch.as_tensor(box, device=dev))


@pytest.mark.parametrize("kind,orthogonal", [("ortho", True), ("triclinic", False), ("open", True)])
def test_pruned_min_dist_equals_brute_force_and_oracle(engine, kind, orthogonal):
    system, rng = _system(kind, 9000, 4.6, 41)
    n = system.n_atoms
    F = 12
    xyz, box, d_xyz, d_box = _frames(engine, system, F)
    if box is not None:                        # per-frame-varying cells (NPT-like)
        scale = (1.0 + 0.01 * np.arange(F, dtype=np.float32))[:, None, None]
        box = (box * scale).astype(np.float32)
        d_box = torch.as_tensor(box, device=engine.device)
    q = np.sort(rng.choice(n, size=700, replace=False))
    h = np.setdiff1d(np.arange(n), q)
    try:
        engine.set_option("prune", 1)
        p0, b0 = engine.get_stat("pruned_frames"), engine.get_stat("brute_frames")
        d1, i1 = engine.min_dist(d_xyz, d_box, q, h, orthogonal)
        torch.cuda.synchronize()
        assert engine.get_stat("pruned_frames") - p0 == F and engine.get_stat("brute_frames") == b0
        engine.set_option("prune", 0)
        d0, i0 = engine.min_dist(d_xyz, d_box, q, h, orthogonal)
    finally:
        engine.set_option("prune", 1)
    assert torch.equal(d1, d0) and torch.equal(i1, i0)
    want_d, want_i = clib.min_dist(xyz[:3], q, h, None if box is None else box[:3], orthogonal)
    assert np.array_equal(d1[:3].cpu().numpy(), want_d) and np.array_equal(i1[:3].cpu().numpy(), want_i)


def test_pruned_min_dist_far_apart_and_mixed_frames(engine):
    """Frames in which no haystack atom is near the query (the stencil minimum is outside the
    guaranteed radius, or there is none) must come from the brute-force kernel; frames of the
    same call that do have a close pair stay pruned."""
    system, rng = _system("ortho", 8000, 6.0, 43)
    n = system.n_atoms
    F = 10
    xyz, box, _, _ = _frames(engine, system, F)
    base = system.base            # (the frames add whole-cell image shifts: irrelevant under the minimum image)
    # query: atoms in a small ball; haystack: everything farther than 1.8 nm from its centre
    centre = np.diag(system.box) / 2
    r = np.linalg.norm(base - centre, axis=1)
    q = np.nonzero(r < 0.8)[0]
    h = np.nonzero(r > 1.8)[0]
    assert len(q) > 10 and len(h) > 4096 and len(q) * len(h) >= (1 << 18)
    # odd frames: one haystack atom is moved right next to a query atom
    for f in range(1, F, 2):
        xyz[f, h[f]] = xyz[f, q[f % len(q)]] + np.float32(0.03)
    d_xyz = torch.as_tensor(xyz, device=engine.device)
    d_box = torch.as_tensor(box, device=engine.device)
    try:
        engine.set_option("prune", 1)
        p0, b0 = engine.get_stat("pruned_frames"), engine.get_stat("brute_frames")
        d1, i1 = engine.min_dist(d_xyz, d_box, q, h, True)
        torch.cuda.synchronize()
        assert engine.get_stat("brute_frames") - b0 == (F + 1) // 2
        assert engine.get_stat("pruned_frames") - p0 == F // 2
        engine.set_option("prune", 0)
        d0, i0 = engine.min_dist(d_xyz, d_box, q, h, True)
    finally:
        engine.set_option("prune", 1)
    assert torch.equal(d1, d0) and torch.equal(i1, i0)
    want_d, want_i = clib.min_dist(xyz[:4], q, h, box[:4], True)
    assert np.array_equal(d1[:4].cpu().numpy(), want_d) and np.array_equal(i1[:4].cpu().numpy(), want_i)


def test_pruned_min_dist_ties_keep_the_first_product_index(engine):
    """numpy argmin returns the FIRST index among equal float32 distances: two haystack atoms at
    exactly the same position next to one query atom, in both list orders."""
    system, rng = _system("ortho", 7000, 5.0, 47)
    n = system.n_atoms
    xyz, box, _, _ = _frames(engine, system, 2)
    q = np.arange(0, 600)
    h = np.arange(600, n)
    xyz[:, h[100]] = xyz[:, q[7]] + np.float32(0.002)
    xyz[:, h[4000]] = xyz[:, h[100]]
    d_xyz = torch.as_tensor(xyz, device=engine.device)
    d_box = torch.as_tensor(box, device=engine.device)
    for hay in (h, h[::-1].copy()):
        d1, i1 = engine.min_dist(d_xyz, d_box, q, hay, True)
        want_d, want_i = clib.min_dist(xyz, q, hay, box, True)
        assert np.array_equal(d1.cpu().numpy(), want_d) and np.array_equal(i1.cpu().numpy(), want_i)
        first = min(int(np.nonzero(hay == h[100])[0][0]), int(np.nonzero(hay == h[4000])[0][0]))
        assert int(i1[0]) == 7 * len(hay) + first


@pytest.mark.parametrize("kind,orthogonal", [("ortho", True), ("triclinic", False), ("open", True)])
@pytest.mark.parametrize("excl_mode", [0, 1, 2])
def test_pruned_nearest_equals_brute_force(engine, kind, orthogonal, excl_mode):
    system, rng = _system(kind, 9000, 4.6, 53)
    n = system.n_atoms
    xyz, box, d_xyz, d_box = _frames(engine, system, 1)
    res, _ = system.topology.atom_residue_chain()
    kwargs = {}
    if excl_mode == 1:
        kwargs["atom_res"] = res
    elif excl_mode == 2:
        ptr = np.arange(0, 3 * n + 1, 3, dtype=np.int64)
        idx = np.stack([np.arange(n), (np.arange(n) + 1) % n, rng.integers(0, n, n)], axis=1).astype(np.int32).reshape(-1)
        kwargs.update(excl_ptr=ptr, excl_idx=idx)
    box9 = None if d_box is None else d_box[0].reshape(9)
    try:
        engine.set_option("prune", 1)
        p0 = engine.get_stat("pruned_frames")
        near1, dist1 = engine.nearest(d_xyz[0], box9, 0.45, orthogonal, excl_mode, **kwargs)
        torch.cuda.synchronize()
        assert engine.get_stat("pruned_frames") == p0 + 1
        engine.set_option("prune", 0)
        near0, dist0 = engine.nearest(d_xyz[0], box9, 0.45, orthogonal, excl_mode, **kwargs)
    finally:
        engine.set_option("prune", 1)
    assert torch.equal(near1, near0) and torch.equal(dist1, dist0)
    assert int((near1 >= 0).sum()) > n // 2
    if excl_mode == 0:      # a sample against the oracle's neighbour list + distances
        neighbours = clib.neighborlist(xyz[0], None if box is None else box[0], 0.45, cells=True)
        for atom in rng.choice(n, size=60, replace=False):
            cand = [int(j) for j in neighbours[atom]]
            if not cand:
                assert int(near1[atom]) == -1
                continue
            d = clib.distances(xyz[:1], np.array([(atom, j) for j in cand], dtype=np.int32),
                               None if box is None else box[:1], orthogonal)[0]
            best = sorted(zip(d.tolist(), cand))[0]
            assert int(near1[atom]) == best[1] and float(dist1[atom]) == np.float32(best[0])


def test_small_periodic_cell_keeps_the_brute_force_kernel(engine):
    """Fewer than three cutoff-sized cells along a periodic dimension: the stencil would alias,
    the pruned search steps aside."""
    system, rng = _system("ortho", 5000, 1.1, 59)          # ~1.1-1.4 nm box: 2-3 cells per side
    xyz, box, d_xyz, d_box = _frames(engine, system, 1)
    b0 = engine.get_stat("brute_frames")
    near1, dist1 = engine.nearest(d_xyz[0], d_box[0].reshape(9), 0.6, True, 0)
    torch.cuda.synchronize()
    assert engine.get_stat("brute_frames") == b0 + 1
    assert int((near1 >= 0).sum()) > 0
