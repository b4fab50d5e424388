"""Timing of the pair-list path on S1 for several batch lengths (build cost = intercept)."""
import json, os, sys, warnings
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from contact_map_b200 import synth
from contact_map_b200.engine import get_engine
warnings.simplefilter("ignore")
engine = get_engine(0); dev = engine.device
which = sys.argv[1] if len(sys.argv) > 1 else "s1"
system = synth.kras_like() if which == "s1" else synth.solvated_triclinic(n_atoms=100000)
frames = [int(a) for a in sys.argv[2:]] or [256, 1024, 4096, 10000, 40000]
res, chain = system.topology.atom_residue_chain()
engine.set_system(res, chain, np.full(system.n_atoms, 3, np.uint8), 0.45, 2)
if os.environ.get("VL_CONSERVATIVE"):
    engine.set_option("verlet_conservative_tiles", 1)
base = torch.as_tensor(system.base, device=dev); grp = torch.as_tensor(system.groups.astype(np.int32), device=dev)
box9 = torch.as_tensor(system.box.reshape(9), device=dev)
a, r = engine.new_tables()
for F in frames:
    xyz = engine.synth_frames(base, grp, box9, system.seed, system.amplitude, True, 0, F)
    box = box9.reshape(1, 3, 3).expand(F, 3, 3).contiguous()
    out = {"frames": F}
    for mode in (0, 2):
        engine.set_option("verlet", mode)
        best = 1e9
        for _ in range(3):
            a.zero_(); r.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            engine.accumulate(xyz, box, a, r)
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out["ms_mode%d" % mode] = best
    out["tripped"] = engine.get_stat("vl_tripped"); out["entries"] = engine.get_stat("vl_entries")
    for k in ("vl_tiles", "vl_tdim", "vl_rows", "vl_slots", "vl_fail", "vl_rebuilds", "vl_rebuild_code", "vl_conservative"):
        out[k[3:]] = engine.get_stat(k)
    print(json.dumps(out), flush=True)
    del xyz, box
