"""e2e step time (public API, pinned host xyz) as a function of the H2D chunk size."""
import os, sys, time, warnings
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import contact_map_b200 as cm
from contact_map_b200 import synth
from contact_map_b200.engine import get_engine
warnings.simplefilter("ignore")
system = synth.kras_like(); top = system.topology; n = top.n_atoms; F = 10000
engine = get_engine(0); dev = engine.device
atom_res, atom_chain = top.atom_residue_chain()
engine.set_system(atom_res, atom_chain, np.full(n, 3, np.uint8), 0.45, 2)
d_xyz = engine.synth_frames(torch.as_tensor(system.base, device=dev), torch.as_tensor(system.groups.astype(np.int32), device=dev),
                            torch.as_tensor(system.box.reshape(9), device=dev), system.seed, system.amplitude, True, 0, F)
h_xyz = torch.empty((F, n, 3), dtype=torch.float32, pin_memory=True); h_xyz.copy_(d_xyz)
h_box = torch.empty((F, 3, 3), dtype=torch.float32, pin_memory=True)
h_box.copy_(torch.as_tensor(system.box).reshape(1, 3, 3).expand(F, 3, 3))
torch.cuda.synchronize()
traj = cm.Trajectory(h_xyz.numpy(), top, h_box.numpy())
everyone = np.arange(n)
for chunk in (0, 10000, 5000, 2500, 1250, 625, 300):
    orig = engine.accumulate_host
    engine.accumulate_host = lambda x, b, a, r, chunk_frames=0, _o=orig, _c=chunk: _o(x, b, a, r, chunk_frames=_c)
    for rep in range(5):
        t0 = time.perf_counter()
        m = cm.ChunkedContactFrequency(traj, query=everyone, haystack=everyone, cutoff=0.45, n_neighbors_ignored=2, presharded=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    engine.accumulate_host = orig
    print("chunk_frames %5d: %.2f ms/step -> %.0f frames/s" % (chunk, dt * 1e3, F / dt))
# raw C call only
atom_t, res_t = engine.new_tables()
for rep in range(3):
    atom_t.zero_(); res_t.zero_(); torch.cuda.synchronize()
    t0 = time.perf_counter(); engine.accumulate_host(h_xyz.numpy(), h_box.numpy(), atom_t, res_t); dt = time.perf_counter() - t0
print("cmb_accumulate_host alone: %.2f ms" % (dt * 1e3))
for mode in (0, 1):
    engine.set_option("verlet", mode)
    for rep in range(3):
        atom_t.zero_(); res_t.zero_(); torch.cuda.synchronize()
        t0 = time.perf_counter(); engine.accumulate_host(h_xyz.numpy(), h_box.numpy(), atom_t, res_t); dt = time.perf_counter() - t0
    print("cmb_accumulate_host alone, verlet=%d: %.2f ms" % (mode, dt * 1e3))
# copy only
d = torch.empty((F, n, 3), dtype=torch.float32, device=dev)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(h_xyz, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("plain H2D of the trajectory: %.2f ms (%.1f GB/s)" % (dt * 1e3, F * n * 12 / dt / 1e9))
