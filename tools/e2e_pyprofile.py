"""cProfile of one e2e step (host side) after warm-up."""
import cProfile, pstats, os, sys, warnings
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
def step():
    m = cm.ChunkedContactFrequency(traj, query=everyone, haystack=everyone, cutoff=0.45, n_neighbors_ignored=2, presharded=True)
    torch.cuda.synchronize()
for _ in range(3): step()
pr = cProfile.Profile(); pr.enable()
for _ in range(5): step()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
