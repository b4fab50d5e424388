"""Host-side stage times of one e2e step (public ChunkedContactFrequency on pinned host xyz):
the same calls as runner.ChunkedContactFrequency._build_contact_map, timestamped."""
import os, sys, time, warnings
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import contact_map_b200 as cm
from contact_map_b200 import synth, parallel
from contact_map_b200.engine import get_engine
from contact_map_b200.contact_map import _engine_system, _split_trajectory, _used_coordinates, _exports_to_pairs, ContactObject
warnings.simplefilter("ignore")
which = sys.argv[1] if len(sys.argv) > 1 else "s1"
system = synth.kras_like() if which == "s1" else synth.solvated_triclinic(n_atoms=100000)
top = system.topology; n = top.n_atoms; F = 10000 if which == "s1" else 4096
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
for _ in range(3):
    m = cm.ChunkedContactFrequency(traj, query=everyone, haystack=everyone, cutoff=0.45, n_neighbors_ignored=2, presharded=True)
acc = {}
def lap(name, t0):
    t = time.perf_counter(); acc[name] = acc.get(name, 0.0) + (t - t0); return t
R = 10 if which == "s1" else 3
for _ in range(R):
    torch.cuda.synchronize()
    t = time.perf_counter()
    obj = ContactObject.__new__(cm.ChunkedContactFrequency)
    ContactObject.__init__(obj, top, everyone, everyone, 0.45, 2)
    t = lap("ContactObject.__init__", t)
    used = _engine_system(obj, engine); t = lap("_engine_system", t)
    xyz, box = _split_trajectory(traj)
    xyz, box, on_device = _used_coordinates(xyz, box, used, top.n_atoms, engine, keep_full_host=True); t = lap("_used_coordinates", t)
    a, r = engine.new_tables(); t = lap("new_tables", t)
    nf = parallel.reduce_frame_count(xyz.shape[0], engine.device); t = lap("reduce_frame_count", t)
    engine.accumulate_host(xyz, box, a, r); t = lap("accumulate_host", t)
    from contact_map_b200.contact_map import _tables_to_pairs
    pairs = _tables_to_pairs(engine, a, r, used); t = lap("_tables_to_pairs", t)
    del a, r, pairs; t = lap("free", t)
    torch.cuda.synchronize(); t = lap("final sync", t)
tot = 0.0
for k, v in acc.items():
    print("%-26s %.3f ms" % (k, v / R * 1e3)); tot += v / R * 1e3
print("%-26s %.3f ms" % ("sum", tot))
t0 = time.perf_counter()
for _ in range(R):
    m = cm.ChunkedContactFrequency(traj, query=everyone, haystack=everyone, cutoff=0.45, n_neighbors_ignored=2, presharded=True)
torch.cuda.synchronize()
print("public call: %.3f ms" % ((time.perf_counter() - t0) / R * 1e3))
