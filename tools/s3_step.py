"""Breakdown of one S3 step (zero tables, accumulate, sparse export) on device-resident frames.

    python tools/s3_step.py [n_frames]
"""
import os, sys, time, warnings
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contact_map_b200 import synth
from contact_map_b200.engine import get_engine
warnings.simplefilter("ignore")
F = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
system = synth.solvated_triclinic(n_atoms=100000)
engine = get_engine(0); dev = engine.device
res, chain = system.topology.atom_residue_chain()
engine.set_system(res, chain, np.full(system.n_atoms, 3, np.uint8), 0.45, 2)
xyz = engine.synth_frames(torch.as_tensor(system.base, device=dev), torch.as_tensor(system.groups.astype(np.int32), device=dev),
                          torch.as_tensor(system.box.reshape(9), device=dev), system.seed, system.amplitude, True, 0, F)
box = torch.as_tensor(system.box, device=dev).reshape(1, 3, 3).expand(F, 3, 3).contiguous()
a, r = engine.new_tables()
def lap(t0):
    torch.cuda.synchronize(); t = time.perf_counter(); return (t - t0) * 1e3, t
for rep in range(4):
    torch.cuda.synchronize(); t = time.perf_counter()
    a.zero_(); r.zero_(); t_zero, t = lap(t)
    engine.accumulate(xyz, box, a, r); t_acc, t = lap(t)
    out = engine.export_many([a, r]); t_exp, t = lap(t)
    print("zero %.2f ms, accumulate %.2f ms, export %.2f ms (%d + %d entries) -> %.0f frames/s"
          % (t_zero, t_acc, t_exp, len(out[0][0]), len(out[1][0]), F / (t_zero + t_acc + t_exp) * 1e3))
