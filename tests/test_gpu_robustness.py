"""Inputs at the edge of the domain: non-finite coordinates must not hang or crash either cell
path (such atoms simply have no contacts; the contacts among the other atoms are unchanged), and
atoms hundreds of nanometres outside the unit cell (unwrapped trajectories) still give exactly
the oracle's counts ("domain")."""
import numpy as np
import pytest
import torch

from contact_map_b200 import synth
from oracle import clib

pytestmark = pytest.mark.gpu

N = 400


def _accumulate(engine, xyz, box):
    atom_t, res_t = engine.new_tables()
    engine.accumulate(torch.as_tensor(xyz, device=engine.device),
                      None if box is None else torch.as_tensor(box, device=engine.device), atom_t, res_t)
    torch.cuda.synchronize()
    return atom_t.view(N, N).cpu().numpy().astype(np.int64), res_t.cpu().numpy().astype(np.int64)


@pytest.mark.parametrize("box_kind", ["ortho", "triclinic", "none"])
@pytest.mark.parametrize("force_path", [-1, 1])
def test_nonfinite_and_far_coordinates(engine, box_kind, force_path):
    rng = np.random.default_rng(5)
    system = synth.random_system(rng, N, n_chains=2, box_kind=box_kind, extent=2.0)
    top = system.topology
    res, chain = (a.astype(np.int32) for a in top.atom_residue_chain())
    role = np.full(N, 3, np.uint8)
    xyz = system.frames(0, 4)
    box = None if system.box is None else np.broadcast_to(system.box, (4, 3, 3)).copy()
    odd_atoms = rng.choice(N, size=12, replace=False)
    engine.set_option("force_path", force_path)
    try:
        engine.set_system(res, chain, role, 0.45, 2)
        clean, _ = _accumulate(engine, xyz, box)

        dirty = xyz.copy()
        for k, a in enumerate(odd_atoms):
            dirty[k % 4, a, k % 3] = [np.nan, np.inf, -np.inf][k % 3]
        got, _ = _accumulate(engine, dirty, box)
        good = np.setdiff1d(np.arange(N), odd_atoms)
        assert np.array_equal(clean[np.ix_(good, good)], got[np.ix_(good, good)])
        assert np.all(got <= clean)

        far = xyz.copy()
        for k, a in enumerate(odd_atoms):
            far[k % 4, a, k % 3] = [300.0, -700.0, 150.0, -420.0][k % 4]
        got_atom, got_res = _accumulate(engine, far, box)
        ref_atom, ref_res = clib.accumulate_dense(far, box, 0.45, res, chain, role, 2, top.n_residues)
        assert np.array_equal(got_atom, ref_atom.astype(np.int64))
        rmap = engine.residue_map
        assert np.array_equal(got_res.reshape(engine.n_res_c, engine.n_res_c),
                              ref_res[np.ix_(rmap, rmap)].astype(np.int64))
    finally:
        engine.set_option("force_path", -1)
