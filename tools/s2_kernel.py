"""Kernel-only time of the per-frame key emission on the S2 systems (ligand / partner vs receptor)."""
import os, sys, warnings
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contact_map_b200 import synth
from contact_map_b200.engine import get_engine
warnings.simplefilter("ignore")
engine = get_engine(0); dev = engine.device
F = 20000
for ligand in (True, False):
    system = synth.binding_system(ligand=ligand)
    top = system.topology; n = system.n_atoms
    res, chain = top.atom_residue_chain()
    role = np.zeros(n, np.uint8); role[system.query] |= 1; role[system.haystack] |= 2
    engine.set_system(res, chain, role, 0.45, 2)
    xyz = engine.synth_frames(torch.as_tensor(system.base, device=dev), torch.as_tensor(system.groups.astype(np.int32), device=dev),
                              torch.as_tensor(system.box.reshape(9), device=dev), system.seed, system.amplitude, system.image_shifts, 0, F)
    box = torch.as_tensor(system.box, device=dev).reshape(1, 3, 3).expand(F, 3, 3).contiguous()
    cap = 8_000_000
    ak = torch.empty(cap, dtype=torch.int64, device=dev); rk = torch.empty(cap, dtype=torch.int64, device=dev)
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        counts = engine._frame_contacts_once(xyz, box, 0, ak, cap, rk, cap, None, None)
        e1.record(); torch.cuda.synchronize()
    a, r = engine.new_tables()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(); engine.accumulate(xyz, box, a, r); e3.record(); torch.cuda.synchronize()
    print("%s: keys kernel %.2f ms (%d atom keys, %d residue keys) -> %.0f frames/s; accumulate %.2f ms -> %.0f frames/s"
          % ("ligand" if ligand else "partner", e0.elapsed_time(e1), counts[0], counts[1], F / e0.elapsed_time(e1) * 1e3,
             e2.elapsed_time(e3), F / e2.elapsed_time(e3) * 1e3))
