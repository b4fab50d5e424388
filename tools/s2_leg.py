import os, sys, time, warnings
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import contact_map_b200 as cm
from contact_map_b200 import synth
from contact_map_b200.engine import get_engine
warnings.simplefilter("ignore")
engine = get_engine(0); dev = engine.device
if len(sys.argv) > 1: engine.set_option("verlet_graph", int(sys.argv[1]))
for ligand in (True, False):
    system = synth.binding_system(ligand=ligand); top = system.topology
    res, chain = top.atom_residue_chain()
    role = np.zeros(system.n_atoms, np.uint8); role[system.query] |= 1; role[system.haystack] |= 2
    engine.set_system(res, chain, role, 0.45, 2)
    F = 20000
    xyz = engine.synth_frames(torch.as_tensor(system.base, device=dev), torch.as_tensor(system.groups.astype(np.int32), device=dev),
                              torch.as_tensor(system.box.reshape(9), device=dev), system.seed, system.amplitude, True, 0, F)
    box = torch.as_tensor(system.box, device=dev).reshape(1, 3, 3).expand(F, 3, 3).contiguous()
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ak, rk = engine.frame_contacts(xyz, box, frame0=0)
        torch.cuda.synchronize(); print("frame_contacts %.1f ms" % ((time.perf_counter() - t0) * 1e3))
    h_xyz = torch.empty((F, system.n_atoms, 3), dtype=torch.float32, pin_memory=True); h_xyz.copy_(xyz)
    h_box = torch.empty((F, 3, 3), dtype=torch.float32, pin_memory=True); h_box.copy_(box)
    torch.cuda.synchronize()
    traj = cm.Trajectory(h_xyz.numpy(), top, h_box.numpy())
    kw = dict(query=system.query, haystack=system.haystack, cutoff=0.45, n_neighbors_ignored=2)
    for rep in range(3):
        t0 = time.perf_counter(); ct = cm.ContactTrajectory(traj, **kw); t1 = time.perf_counter()
        a = cm.ContactFrequency(traj[:F // 2], **kw); t2 = time.perf_counter()
        b = cm.ContactFrequency(traj[F // 2:], **kw); t3 = time.perf_counter()
        d = a - b; n = len(d.atom_contacts.counter); t4 = time.perf_counter()
        print("ligand=%s CT %.1f ms, CF %.1f + %.1f ms, diff %.1f ms, graph_state %d" % (ligand, (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3, engine.get_stat("vl_graph_state")))
