"""GPU check of the frame-batch pair-list path (cmb_verlet.cuh) against the cell kernels and the
oracle, with list statistics and timings.  python tools/verlet_check.py [--s3] > gpurun_out/verlet.log"""
import json
import os
import sys
import time
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from contact_map_b200 import synth                     # noqa: E402
from contact_map_b200.engine import get_engine         # noqa: E402
from oracle import clib                                # noqa: E402

warnings.simplefilter("ignore")
engine = get_engine(0)
dev = engine.device
STATS = ("vl_used", "vl_valid", "vl_fail", "vl_tripped", "vl_entries", "vl_tiles", "vl_rows", "vl_slots", "vl_segs",
         "vl_bundles", "vl_rlist_um")


def stats():
    return {k: engine.get_stat(k) for k in STATS}


def device_frames(system, frame0, n_frames):
    xyz = engine.synth_frames(torch.as_tensor(system.base, device=dev),
                              torch.as_tensor(system.groups.astype(np.int32), device=dev),
                              None if system.box is None else torch.as_tensor(system.box.reshape(9), device=dev),
                              system.seed, system.amplitude, system.image_shifts and system.box is not None, frame0, n_frames)
    box = None if system.box is None else torch.as_tensor(system.box, device=dev).reshape(1, 3, 3).expand(n_frames, 3, 3).contiguous()
    return xyz, box


def roles(system):
    n = system.n_atoms
    role = np.zeros(n, dtype=np.uint8)
    role[np.arange(n) if system.query is None else system.query] |= 1
    role[np.arange(n) if system.haystack is None else system.haystack] |= 2
    return role


def run(xyz, box, mode, audit=0):
    engine.set_option("verlet", mode)
    engine.set_option("verlet_audit", audit)
    a, r = engine.new_tables()
    engine.accumulate(xyz, box, a, r)
    torch.cuda.synchronize()
    return a, r


def timed(xyz, box, mode, rep=3):
    engine.set_option("verlet", mode)
    engine.set_option("verlet_audit", 0)
    a, r = engine.new_tables()
    best = 1e9
    for _ in range(rep):
        a.zero_(); r.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        engine.accumulate(xyz, box, a, r)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def check_system(name, system, n_frames, oracle_frames, n_ignored=2, cutoff=0.45, perturb=None):
    top = system.topology
    res, chain = top.atom_residue_chain()
    role = roles(system)
    engine.set_system(res, chain, role, cutoff, n_ignored)
    xyz, box = device_frames(system, 0, n_frames)
    if perturb is not None:
        xyz = perturb(xyz)
    a0, r0 = run(xyz, box, 0)
    a1, r1 = run(xyz, box, 2)
    st = stats()
    same = bool(torch.equal(a0, a1) and torch.equal(r0, r1))
    a2, r2 = run(xyz, box, 2, audit=1)
    audited, disagree = engine.get_stat("vl_audited"), engine.get_stat("vl_disagree")
    same_audit = bool(torch.equal(a0, a2) and torch.equal(r0, r2))
    out = {"name": name, "n": system.n_atoms, "frames": n_frames, "verlet_equals_cells": same,
           "audit_equals_cells": same_audit, "audited": audited, "disagree": disagree, "stats": st,
           "nonzero_atom": int(torch.count_nonzero(a0)), "nonzero_res": int(torch.count_nonzero(r0))}
    if oracle_frames:
        ao, ro = run(xyz[:oracle_frames].contiguous(), None if box is None else box[:oracle_frames].contiguous(), 2)
        st2 = stats()
        n = system.n_atoms
        xh = xyz[:oracle_frames].cpu().numpy()
        bh = None if box is None else box[:oracle_frames].cpu().numpy()
        want_a, want_r = clib.accumulate_dense(xh, bh, cutoff, res, chain, role, n_ignored, engine.n_res_c if False else int(res.max()) + 1)
        got_a = ao.cpu().numpy().view(np.uint32).reshape(n, n)
        rc = engine.n_res_c
        got_r = ro.cpu().numpy().view(np.uint32).reshape(rc, rc)
        rmap = engine.residue_map
        full_r = np.zeros_like(want_r)
        full_r[np.ix_(rmap, rmap)] = got_r
        out["oracle_frames"] = oracle_frames
        out["oracle_atom_equal"] = bool(np.array_equal(got_a, want_a))
        out["oracle_res_equal"] = bool(np.array_equal(full_r, want_r))
        out["oracle_stats"] = st2
    print(json.dumps(out), flush=True)
    return out


def check_big(name, system, n_frames, cutoff=0.45, n_ignored=2):
    """Large systems (dense tables of tens of GB): one pair of tables at a time, compared through
    their sparse exports."""
    top = system.topology
    res, chain = top.atom_residue_chain()
    engine.set_system(res, chain, roles(system), cutoff, n_ignored)
    xyz, box = device_frames(system, 0, n_frames)
    a, r = engine.new_tables()
    exports, st, audit = [], None, None
    for mode, aud in ((0, 0), (2, 0), (2, 1)):
        a.zero_(); r.zero_()
        engine.set_option("verlet", mode)
        engine.set_option("verlet_audit", aud)
        engine.accumulate(xyz, box, a, r)
        torch.cuda.synchronize()
        if mode == 2 and not aud:
            st = stats()
        if aud:
            audit = (engine.get_stat("vl_audited"), engine.get_stat("vl_disagree"))
        exports.append(engine.export_many([a, r]))
    def same(x, y):
        return all(np.array_equal(p[0], q[0]) and np.array_equal(p[1], q[1]) for p, q in zip(x, y))
    out = {"name": name, "n": system.n_atoms, "frames": n_frames, "verlet_equals_cells": same(exports[0], exports[1]),
           "audit_equals_cells": same(exports[0], exports[2]), "audited": audit[0], "disagree": audit[1], "stats": st,
           "nonzero_atom": int(len(exports[0][0][0])), "nonzero_res": int(len(exports[0][1][0]))}
    print(json.dumps(out), flush=True)
    del a, r
    torch.cuda.empty_cache()
    return out


def main():
    ok = True
    rng = np.random.default_rng(11)
    if "--only-s3" in sys.argv:
        s3 = synth.solvated_triclinic(n_atoms=100000)
        o = check_big("S3", s3, 64)
        ok &= o["verlet_equals_cells"] and o["audit_equals_cells"] and o["disagree"] == 0
        res, chain = s3.topology.atom_residue_chain()
        engine.set_system(res, chain, roles(s3), 0.45, 2)
        xyz, box = device_frames(s3, 0, 256)
        t = {"S3_256_ms_cells": timed(xyz, box, 0, rep=2), "S3_256_ms_verlet": timed(xyz, box, 2, rep=2)}
        t["stats"] = stats()
        print(json.dumps(t), flush=True)
        mid = synth.solvated_triclinic(n_atoms=20000, lattice=28)
        o = check_big("mid-20k", mid, 128)
        ok &= o["verlet_equals_cells"] and o["audit_equals_cells"] and o["disagree"] == 0
        print("VERLET_CHECK", "OK" if ok else "FAILED", flush=True)
        return 0 if ok else 1
    # S1
    s1 = synth.kras_like()
    o = check_system("S1", s1, 1024, 32)
    ok &= o["verlet_equals_cells"] and o["audit_equals_cells"] and o["disagree"] == 0 and o.get("oracle_atom_equal", True) and o.get("oracle_res_equal", True)
    # S2 ligand / partner
    for lig in (True, False):
        o = check_system("S2-ligand" if lig else "S2-partner", synth.binding_system(ligand=lig), 512, 16)
        ok &= o["verlet_equals_cells"] and o["audit_equals_cells"] and o["disagree"] == 0 and o["oracle_atom_equal"] and o["oracle_res_equal"]
    # random systems: ortho / triclinic / open, several chains, hydrogens, different n_ignored
    for k, (kind, n, ext) in enumerate((("ortho", 900, 2.6), ("triclinic", 1200, 2.8), ("open", 700, 2.2),
                                        ("triclinic", 3000, 3.6), ("ortho", 300, 1.5))):
        sysm = synth.random_system(rng, n, n_chains=3, box_kind=kind, extent=ext)
        o = check_system("random-%s-%d" % (kind, n), sysm, 300, 24, n_ignored=k % 4)
        ok &= o["verlet_equals_cells"] and o["audit_equals_cells"] and o["disagree"] == 0 and o["oracle_atom_equal"] and o["oracle_res_equal"]
    # guard trips: some frames get a few atoms kicked far away / a frame fully scrambled
    def kick(xyz):
        xyz = xyz.clone()
        xyz[100, 5] += 0.4
        xyz[101, 700:720] -= 0.3
        xyz[400] = xyz[400].flip(0)
        xyz[777, 2000, 0] = float("nan")
        return xyz
    o = check_system("S1-kicked", s1, 1024, 0, perturb=kick)
    ok &= o["verlet_equals_cells"] and o["audit_equals_cells"] and o["disagree"] == 0 and o["stats"]["vl_tripped"] >= 4
    # timings S1
    res, chain = s1.topology.atom_residue_chain()
    engine.set_system(res, chain, roles(s1), 0.45, 2)
    xyz, box = device_frames(s1, 0, 10000)
    t = {"S1_10k_ms_cells": timed(xyz, box, 0), "S1_10k_ms_verlet": timed(xyz, box, 2)}
    t["stats"] = stats()
    print(json.dumps(t), flush=True)
    if "--s3" in sys.argv:
        s3 = synth.solvated_triclinic(n_atoms=100000)
        o = check_big("S3", s3, 64)
        ok &= o["verlet_equals_cells"] and o["audit_equals_cells"] and o["disagree"] == 0
        res, chain = s3.topology.atom_residue_chain()
        engine.set_system(res, chain, roles(s3), 0.45, 2)
        xyz, box = device_frames(s3, 0, 256)
        t = {"S3_256_ms_cells": timed(xyz, box, 0, rep=2), "S3_256_ms_verlet": timed(xyz, box, 2, rep=2)}
        t["stats"] = stats()
        print(json.dumps(t), flush=True)
    print("VERLET_CHECK", "OK" if ok else "FAILED", flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
