"""Timeline of one streamed host call (option stage_trace): when the copy and the kernels of every
chunk finish, relative to the first copy being issued.

    python tools/e2e_trace.py [s1|s3] [chunk_frames ...]
"""
import os, sys, time, warnings
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contact_map_b200 import synth
from contact_map_b200.engine import get_engine
warnings.simplefilter("ignore")
which = "s1"
if len(sys.argv) > 1 and sys.argv[1] in ("s1", "s3"):
    which = sys.argv.pop(1)
system = synth.kras_like() if which == "s1" else synth.solvated_triclinic(n_atoms=100000)
top = system.topology; n = top.n_atoms; F = 10000 if which == "s1" else 2048
engine = get_engine(0); dev = engine.device
atom_res, atom_chain = top.atom_residue_chain()
engine.set_system(atom_res, atom_chain, np.full(n, 3, np.uint8), 0.45, 2)
d_xyz = engine.synth_frames(torch.as_tensor(system.base, device=dev), torch.as_tensor(system.groups.astype(np.int32), device=dev),
                            torch.as_tensor(system.box.reshape(9), device=dev), system.seed, system.amplitude, True, 0, F)
h_xyz = torch.empty((F, n, 3), dtype=torch.float32, pin_memory=True); h_xyz.copy_(d_xyz)
h_box = torch.empty((F, 3, 3), dtype=torch.float32, pin_memory=True)
h_box.copy_(torch.as_tensor(system.box).reshape(1, 3, 3).expand(F, 3, 3))
torch.cuda.synchronize()
atom_t, res_t = engine.new_tables()
engine.set_option("stage_trace", 1)
engine.set_option("time_kernels", 1)
for chunk in [int(a) for a in sys.argv[1:]] or [0, 5000, 1250]:
    for rep in range(4):
        atom_t.zero_(); res_t.zero_(); torch.cuda.synchronize()
        engine.get_stat("vl_frames_ns")
        t0 = time.perf_counter()
        engine.accumulate_host(h_xyz.numpy(), h_box.numpy(), atom_t, res_t, chunk_frames=chunk)
        dt = time.perf_counter() - t0
    k = engine.get_stat("trace_chunks")
    print("chunk_frames %d: host call %.3f ms, frame kernels %.3f ms in %d launches" %
          (chunk, dt * 1e3, engine.get_stat("vl_frames_ns") * 1e-6, k))
    for c in (range(k) if k <= 12 else list(range(6)) + list(range(k - 4, k))):
        print("   chunk %d: copy issued+done at %.3f ms, kernels done at %.3f ms" %
              (c, engine.get_stat("trace_copy_%d" % c) * 1e-6, engine.get_stat("trace_done_%d" % c) * 1e-6))
