"""GPU parity fuzz: seeded random systems, boxes that change from frame to frame (NPT-like),
cutoffs from well below to above a third of the box (the latter forces the exact-all fallback of
the screening margins), random roles and exclusion widths, both cell paths -- bit-exact against
the CPU oracle (checker only)."""
import numpy as np
import pytest
import torch

from contact_map_b200 import synth
from oracle import clib
from tests.helpers import system_arrays, dense_from_export

pytestmark = pytest.mark.gpu

N_CASES = 36


def _case(seed):
    rng = np.random.default_rng(9000 + seed)
    n = int(rng.choice([2, 5, 17, 64, 130, 333, 700, 1100]))
    box_kind = ["ortho", "triclinic", "none"][seed % 3]
    extent = float(rng.uniform(1.0, 3.2)) if n > 17 else float(rng.uniform(0.6, 1.1))
    system = synth.random_system(rng, n, n_chains=int(rng.integers(1, 5)), box_kind=box_kind, extent=extent)
    n_frames = int(rng.integers(2, 7))
    xyz = system.frames(0, n_frames)
    box = None
    if system.box is not None:
        # independent per-frame scaling of the cell rows (and of the tilt for triclinic cells);
        # coordinates are left alone: MDTraj's minimum image handles atoms outside the cell
        scale = rng.uniform(0.94, 1.07, size=(n_frames, 3, 1)).astype(np.float32)
        box = (system.box[None, :, :] * scale).astype(np.float32)
        if seed % 6 == 1:
            box[1:] = box[0]           # constant after the first frame: the cached-grid branch
    cutoff = float(rng.choice([0.25, 0.33, 0.45, 0.45, 0.6, 0.8, 0.37 * extent]))
    n_ign = int(rng.integers(0, 5))
    qf, hf = rng.choice([0.15, 0.5, 1.0]), rng.choice([0.3, 1.0, 1.0])
    query = np.sort(rng.choice(n, size=max(1, int(qf * n)), replace=False))
    hay = np.sort(rng.choice(n, size=max(1, int(hf * n)), replace=False))
    force = 1 if seed % 4 == 3 else -1
    return system, xyz, box, cutoff, n_ign, query, hay, force


@pytest.mark.parametrize("seed", range(N_CASES))
def test_random_configuration_matches_oracle(engine, seed):
    system, xyz, box, cutoff, n_ign, query, hay, force = _case(seed)
    n = system.topology.n_atoms
    res, chain, role = system_arrays(system, query, hay)
    ref_atom, ref_res = clib.accumulate_dense(xyz, box, cutoff, res, chain, role, n_ign,
                                              system.topology.n_residues)
    engine.set_option("force_path", force)
    # global-memory cell path: alternate between its two schedules (one frame at a time / batched)
    engine.set_option("large_batch", 0 if (seed // 4) % 2 else 2)
    try:
        engine.set_system(res, chain, role, cutoff, n_ign)
        atom_t, res_t = engine.new_tables()
        d_xyz = torch.as_tensor(xyz, device=engine.device)
        d_box = None if box is None else torch.as_tensor(box, device=engine.device)
        engine.accumulate(d_xyz, d_box, atom_t, res_t)
        idx, cnt = engine.export_nonzero(atom_t)
        assert np.array_equal(dense_from_export(idx, cnt, n), ref_atom)
        ridx, rcnt = engine.export_nonzero(res_t)
        rmap = engine.residue_map
        assert np.array_equal(dense_from_export(ridx, rcnt, engine.n_res_c), ref_res[np.ix_(rmap, rmap)])
        assert int(ref_res.sum()) == int(ref_res[np.ix_(rmap, rmap)].sum())
        # the per-frame key stream describes the same contacts
        akeys, rkeys = engine.frame_contacts(d_xyz, d_box)
        assert len(akeys) == int(ref_atom.sum()) and len(rkeys) == int(ref_res.sum())
    finally:
        engine.set_option("force_path", -1)
        engine.set_option("large_batch", 1)
