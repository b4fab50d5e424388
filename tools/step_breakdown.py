"""Host/device time breakdown of one bench step (inputs resident in HBM).
    python tools/step_breakdown.py [s1|s3] [frames]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contact_map_b200 import synth
from contact_map_b200.engine import get_engine
which = sys.argv[1] if len(sys.argv) > 1 else 's1'
system = synth.kras_like() if which == 's1' else synth.solvated_triclinic(n_atoms=100000)
top = system.topology; n = top.n_atoms; F = int(sys.argv[2]) if len(sys.argv) > 2 else (10000 if which == 's1' else 1024)
engine = get_engine(0); dev = engine.device
atom_res, atom_chain = top.atom_residue_chain()
engine.set_system(atom_res, atom_chain, np.full(n, 3, np.uint8), 0.45, 2)
d_xyz = engine.synth_frames(torch.as_tensor(system.base, device=dev), torch.as_tensor(system.groups.astype(np.int32), device=dev),
                            torch.as_tensor(system.box.reshape(9), device=dev), system.seed, system.amplitude, True, 0, F)
d_box = torch.as_tensor(system.box, device=dev).reshape(1, 3, 3).expand(F, 3, 3).contiguous()
atom_t, res_t = engine.new_tables()
def sync(): torch.cuda.synchronize()
for rep in range(6):
    sync(); t = [time.perf_counter()]
    atom_t.zero_(); res_t.zero_(); sync(); t.append(time.perf_counter())
    engine.accumulate(d_xyz, d_box, atom_t, res_t); sync(); t.append(time.perf_counter())
    out = engine.export_many([atom_t, res_t]); t.append(time.perf_counter())
    if rep >= 3:
        print("zero %.0f us, kernel %.0f us, export %.0f us" % tuple(1e6 * (b - a) for a, b in zip(t, t[1:])))
