"""Wall-clock of the whole per-frame key pipeline (emit -> device sort -> D2H) on the S2 ligand system."""
import os, sys, time, warnings
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contact_map_b200 import synth
from contact_map_b200.engine import get_engine
warnings.simplefilter("ignore")
engine = get_engine(0); dev = engine.device
F = 100000
system = synth.binding_system(ligand=True)
top = system.topology; n = system.n_atoms
res, chain = top.atom_residue_chain()
role = np.zeros(n, np.uint8); role[system.query] |= 1; role[system.haystack] |= 2
engine.set_system(res, chain, role, 0.45, 2)
xyz = engine.synth_frames(torch.as_tensor(system.base, device=dev), torch.as_tensor(system.groups.astype(np.int32), device=dev),
                          torch.as_tensor(system.box.reshape(9), device=dev), system.seed, system.amplitude, system.image_shifts, 0, F)
box = torch.as_tensor(system.box, device=dev).reshape(1, 3, 3).expand(F, 3, 3).contiguous()
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ak, rk = engine.frame_contacts(xyz, box)
    dt = time.perf_counter() - t0
    print("call %d: %.1f ms -> %.0f frames/s (%d atom keys, %d residue keys)" % (rep, dt * 1e3, F / dt, len(ak), len(rk)))
