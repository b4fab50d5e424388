"""Small end-to-end run of every kernel variant (also the input for compute-sanitizer where the pool
allows it): both paths,
ortho / triclinic / no box, roles, per-frame keys, hashed tables."""
import os, sys, warnings
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contact_map_b200 import synth
from contact_map_b200.engine import get_engine, HashTable
warnings.simplefilter("ignore")
engine = get_engine(0)
rng = np.random.default_rng(1)
for box_kind in ("ortho", "triclinic", "none"):
    system = synth.random_system(rng, 500, n_chains=2, box_kind=box_kind, extent=2.2)
    top = system.topology
    res, chain = top.atom_residue_chain()
    role = np.full(500, 3, np.uint8); role[::3] = 1; role[1::3] = 2
    xyz = torch.as_tensor(system.frames(0, 3), device=engine.device)
    box = None if system.box is None else torch.as_tensor(np.broadcast_to(system.box, (3, 3, 3)).copy(), device=engine.device)
    for force in (0, 1):
        engine.set_option("force_path", force if force else -1)
        engine.set_system(res, chain, role, 0.45, 2)
        a, r = engine.new_tables(hashed=False)
        engine.accumulate(xyz, box, a, r)
        ak, rk = engine.frame_contacts(xyz, box)
        engine.boundary_pairs(xyz, box, ulps=64)
        if force:
            ha, hr = HashTable(12, engine.device), HashTable(8, engine.device)
            engine.accumulate(xyz, box, ha, hr)
            out = engine.export_many([ha, hr])
        out = engine.export_many([a, r])
        print(box_kind, force, int(out[0][1].sum()), len(ak), len(rk))
engine.set_option("force_path", -1)
torch.cuda.synchronize()
print("done")
