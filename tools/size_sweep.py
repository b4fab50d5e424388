"""Frame kernel throughput against system size (globular protein of n_res x 10 heavy atoms):
where the fused kernel drops from two CTAs per SM to one, and where the global-memory path takes over."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contact_map_b200 import synth
from contact_map_b200.engine import get_engine
engine = get_engine(0); dev = engine.device
for n_res in [int(a) for a in sys.argv[1:]] or (100, 200, 270, 280, 300, 400, 500, 614, 700, 1000, 2000, 5000):
    edge = 7.0 * (n_res / 270.0) ** (1.0 / 3.0)
    system = synth.kras_like(n_res=n_res, box=(edge, edge * 1.07, edge * 1.14))
    top = system.topology; n = top.n_atoms
    res, chain = top.atom_residue_chain()
    engine.set_system(res, chain, np.full(n, 3, np.uint8), 0.45, 2)
    F = max(256, int(2.7e7 / n) // 296 * 296)
    xyz = engine.synth_frames(torch.as_tensor(system.base, device=dev), torch.as_tensor(system.groups.astype(np.int32), device=dev),
                              torch.as_tensor(system.box.reshape(9), device=dev), system.seed, system.amplitude, True, 0, F)
    box = torch.as_tensor(system.box, device=dev).reshape(1, 3, 3).expand(F, 3, 3).contiguous()
    a, r = engine.new_tables()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); engine.accumulate(xyz, box, a, r); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("n = %5d atoms, path %d: %8.0f frames/s, %6.2f M atom-frames/s" % (n, engine.path_kind, F / best * 1e3, n * F / best / 1e3))
