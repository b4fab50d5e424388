"""e2e step time of the public API with PAGEABLE (ordinary numpy) vs pinned host coordinates."""
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
h_pinned = torch.empty((F, n, 3), dtype=torch.float32, pin_memory=True); h_pinned.copy_(d_xyz)
box = np.broadcast_to(system.box, (F, 3, 3)).astype(np.float32).copy()
pageable = h_pinned.numpy().copy()
want = None
for name, xyz, staged in (("pinned", h_pinned.numpy(), 1), ("pageable staged", pageable, 1), ("pageable direct", pageable, 0)):
    engine.set_option("stage_pageable", staged)
    traj = cm.Trajectory(xyz, top, box)
    for rep in range(5):
        t0 = time.perf_counter()
        m = cm.ContactFrequency(traj, cutoff=0.45, n_neighbors_ignored=2)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print("%-16s %.2f ms/step -> %.0f frames/s" % (name, dt * 1e3, F / dt))
    if want is None:
        want = m._atom_contacts
    assert m._atom_contacts == want
