"""cProfile of the public ContactTrajectory / ContactFrequency API on the S2 system from pinned host
memory (after warm-up):  python tools/s2_api_profile.py [ligand|partner] [frames] [traj|freq]"""
import cProfile, pstats, os, sys, time, warnings
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import contact_map_b200 as cm
from contact_map_b200 import synth
from contact_map_b200.engine import get_engine
warnings.simplefilter("ignore")
ligand = (sys.argv[1] if len(sys.argv) > 1 else "ligand") == "ligand"
F = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
what = sys.argv[3] if len(sys.argv) > 3 else "traj"
system = synth.kras_like() if what == "s1" else synth.binding_system(ligand=ligand)
top = system.topology; n = top.n_atoms
engine = get_engine(0); dev = engine.device
d_xyz = engine.synth_frames(torch.as_tensor(system.base, device=dev), torch.as_tensor(system.groups.astype(np.int32), device=dev),
                            torch.as_tensor(system.box.reshape(9), device=dev), system.seed, system.amplitude, True, 0, F)
h_xyz = torch.empty((F, n, 3), dtype=torch.float32, pin_memory=True); h_xyz.copy_(d_xyz)
h_box = torch.empty((F, 3, 3), dtype=torch.float32, pin_memory=True)
h_box.copy_(torch.as_tensor(system.box).reshape(1, 3, 3).expand(F, 3, 3))
torch.cuda.synchronize(); del d_xyz
traj = cm.Trajectory(h_xyz.numpy(), top, h_box.numpy())
q = np.arange(n) if system.query is None else system.query
h = np.arange(n) if system.haystack is None else system.haystack
kw = dict(query=q, haystack=h, cutoff=0.45, n_neighbors_ignored=2)
def step():
    if what == "traj":
        return cm.ContactTrajectory(traj, **kw)
    return cm.ContactFrequency(traj, **kw)
for _ in range(2): step()
t0 = time.perf_counter()
for _ in range(3): step()
dt = (time.perf_counter() - t0) / 3
print("%s %s: %d frames in %.2f ms = %.0f frames/s (H2D floor %.2f ms at 55 GB/s)" % (
    "ligand" if ligand else "partner", what, F, dt * 1e3, F / dt, F * n * 12 / 55e9 * 1e3))
print({k: round(v * 1e3, 2) for k, v in getattr(engine, "last_timing", {}).items()})
t0 = time.perf_counter(); x = torch.empty(20000000, dtype=torch.int32, pin_memory=True); t1 = time.perf_counter(); del x
y = torch.empty(20000000, dtype=torch.int32, pin_memory=True); t2 = time.perf_counter()
print("pinned 80 MB alloc: first %.2f ms, after free %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
pr = cProfile.Profile(); pr.enable()
for _ in range(3): step()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
