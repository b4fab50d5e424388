"""S3 (100 000 atoms, 32 704 residues): the batched mode (per-frame residue bitmaps) and the
one-frame-at-a-time mode (frame stamps) of the global-memory cell path give identical tables."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contact_map_b200 import synth
from contact_map_b200.engine import get_engine
engine = get_engine(0); dev = engine.device
system = synth.solvated_triclinic(n_atoms=100000)
top = system.topology; n = top.n_atoms
res, chain = top.atom_residue_chain()
F = 13
xyz = engine.synth_frames(torch.as_tensor(system.base, device=dev), torch.as_tensor(system.groups.astype(np.int32), device=dev),
                          torch.as_tensor(system.box.reshape(9), device=dev), system.seed, system.amplitude, True, 0, F)
box = torch.as_tensor(system.box, device=dev).reshape(1, 3, 3).expand(F, 3, 3).contiguous()
out = []
for mode in (0, 2):
    engine.set_option("large_batch", mode)
    engine.set_system(res, chain, np.full(n, 3, np.uint8), 0.45, 2)
    a, r = engine.new_tables()
    engine.accumulate(xyz, box, a, r)
    (ai, ac), (ri, rc) = engine.export_many([a, r])
    out.append((ai, ac, ri, rc))
    print("mode %d: %d atom pairs (sum %d), %d residue pairs (sum %d)" % (mode, len(ai), ac.sum(), len(ri), rc.sum()))
    del a, r
    torch.cuda.empty_cache()
same = all(np.array_equal(x, y) for x, y in zip(out[0], out[1]))
print("identical:", same)
sys.exit(0 if same else 1)
