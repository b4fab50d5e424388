import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
from contact_map_b200 import synth
from contact_map_b200.engine import get_engine, HashTable
system = synth.solvated_triclinic(n_atoms=100000)
top = system.topology; n = top.n_atoms
engine = get_engine(0)
atom_res, atom_chain = top.atom_residue_chain()
role = np.full(n, 3, dtype=np.uint8)
engine.set_system(atom_res, atom_chain, role, 0.45, 2)
dev = engine.device
F = 64
d_xyz = engine.synth_frames(torch.as_tensor(system.base, device=dev), torch.as_tensor(system.groups.astype(np.int32), device=dev),
                            torch.as_tensor(system.box.reshape(9), device=dev), system.seed, system.amplitude, True, 0, F)
d_box = torch.as_tensor(system.box, device=dev).reshape(1,3,3).expand(F,3,3).contiguous()
def run(name, mk):
    for rep in range(3):
        a, r = mk()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); engine.accumulate(d_xyz, d_box, a, r); e1.record(); torch.cuda.synchronize()
        print(name, "%.1f us/frame" % (e0.elapsed_time(e1)*1e3/F))
        del a, r
rc = engine.n_res_c
run("dense/dense", lambda: engine.new_tables(hashed=False))
run("dense/hashed-res", lambda: (engine.new_tables(hashed=False)[0], HashTable(21, dev)))
run("hashed/hashed", lambda: engine.new_tables(hashed=True))
run("hashed/none", lambda: (engine.new_tables(hashed=True)[0], None))
run("dense/none", lambda: (engine.new_tables(hashed=False)[0], None))
