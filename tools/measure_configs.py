"""Measure the BASELINE.json configs other than the bench headline (S1) on one GPU.

    python tools/measure_configs.py [--quick] > gpurun_out/configs.json

S2: ContactTrajectory per-frame output (ligand / partner chain vs receptor) + ContactDifference
S3: ContactFrequency on the 100 000-atom triclinic system (global-memory cell path)
S4: MinimumDistanceCounter / NearestAtoms on S3, chunked runner on S1
Every number is device work timed with CUDA events (inputs resident in HBM) unless it says e2e.
"""
import json
import os
import sys
import time
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import contact_map_b200 as cm                          # noqa: E402
from contact_map_b200 import synth                     # noqa: E402
from contact_map_b200.engine import get_engine         # noqa: E402

warnings.simplefilter("ignore", PendingDeprecationWarning)
PEAK_HBM = 6554.6
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK_HBM = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])


def device_frames(engine, system, frame0, n_frames):
    dev = engine.device
    xyz = engine.synth_frames(torch.as_tensor(system.base, device=dev),
                              torch.as_tensor(system.groups.astype(np.int32), device=dev),
                              torch.as_tensor(system.box.reshape(9), device=dev), system.seed,
                              system.amplitude, system.image_shifts, frame0, n_frames)
    box = torch.as_tensor(system.box, device=dev).reshape(1, 3, 3).expand(n_frames, 3, 3).contiguous()
    return xyz, box


def roles(system):
    n = system.n_atoms
    role = np.zeros(n, dtype=np.uint8)
    role[np.arange(n) if system.query is None else system.query] |= 1
    role[np.arange(n) if system.haystack is None else system.haystack] |= 2
    return role


def timed(fn, repeat=3):
    best = None
    for _ in range(repeat):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return best, out


class DevTraj(object):
    def __init__(self, topology, xyz, box):
        self.topology, self.xyz, self.unitcell_vectors = topology, xyz, box

    def __len__(self):
        return self.xyz.shape[0]

    def __getitem__(self, key):
        return DevTraj(self.topology, self.xyz[key], self.unitcell_vectors[key])


def main():
    quick = "--quick" in sys.argv
    engine = get_engine(0)
    out = {"gpu": torch.cuda.get_device_name(0), "hbm_peak_gbs": PEAK_HBM}

    # ---- S2 -----------------------------------------------------------------------------
    for ligand in (True, False):
        system = synth.binding_system(ligand=ligand)
        top = system.topology
        F = 20000 if quick else 100000
        res, chain = top.atom_residue_chain()
        engine.set_system(res, chain, roles(system), 0.45, 2)
        chunk = 20000
        passes = []
        for rep in range(2):          # pass 0 also pays for the pinned key buffers, pass 1 is steady state
            n_keys = 0
            t_total = 0.0
            for f0 in range(0, F, chunk):
                xyz, box = device_frames(engine, system, f0, chunk)
                ms, (ak, rk) = timed(lambda: engine.frame_contacts(xyz, box, frame0=0), repeat=1)
                t_total += ms
                n_keys += len(ak)
                del ak, rk
            passes.append(t_total)
        name = "S2_ligand" if ligand else "S2_partner"
        b_alg = F * (12 * system.n_atoms + 36) + 8 * n_keys
        out[name + "_contact_trajectory"] = {
            "frames": F, "n_atoms": system.n_atoms, "atom_keys": n_keys, "ms": t_total, "first_pass_ms": passes[0],
            "frames_per_s": F / t_total * 1e3, "includes": "key emission kernels, radix sort, D2H of the keys",
            "hbm_algorithmic_gbs": b_alg / t_total / 1e6, "hbm_frac": b_alg / t_total / 1e6 / PEAK_HBM}
        # public API: ContactTrajectory + ContactDifference(first half, second half)
        Fa = 4000 if quick else 20000
        xyz, box = device_frames(engine, system, 0, Fa)
        traj = DevTraj(top, xyz, box)
        t0 = time.perf_counter()
        ct = cm.ContactTrajectory(traj, query=system.query, haystack=system.haystack)
        t1 = time.perf_counter()
        a = cm.ContactFrequency(traj[:Fa // 2], query=system.query, haystack=system.haystack)
        b = cm.ContactFrequency(traj[Fa // 2:], query=system.query, haystack=system.haystack)
        diff = a - b
        n_diff = len(diff.atom_contacts.counter)
        t2 = time.perf_counter()
        out[name + "_api"] = {"frames": Fa, "ContactTrajectory_s": t1 - t0,
                              "ContactTrajectory_frames_per_s": Fa / (t1 - t0),
                              "two_ContactFrequency_plus_ContactDifference_s": t2 - t1,
                              "difference_pairs": n_diff, "frames_in_trajectory": len(ct)}

    # ---- S3 -----------------------------------------------------------------------------
    system = synth.solvated_triclinic(n_atoms=100000)
    top = system.topology
    res, chain = top.atom_residue_chain()
    engine.set_system(res, chain, roles(system), 0.45, 2)
    F3 = 64 if quick else 256
    xyz, box = device_frames(engine, system, 0, F3)
    atom_t, res_t = engine.new_tables()
    engine.accumulate(xyz[:8], box[:8], atom_t, res_t)           # page in the tables
    ms, _ = timed(lambda: engine.accumulate(xyz, box, atom_t, res_t))
    engine.set_option("count_tests", 1)
    engine.accumulate(xyz[:4], box[:4], atom_t, res_t)
    tests_per_frame = engine.get_stat("pair_tests") / 4
    engine.set_option("count_tests", 0)
    b_alg = F3 * (12 * system.n_atoms + 36)
    out["S3_contact_frequency"] = {
        "frames": F3, "n_atoms": system.n_atoms, "n_residues": top.n_residues, "ms": ms,
        "frames_per_s": F3 / ms * 1e3, "pair_tests_per_frame": tests_per_frame,
        "pair_tests_per_s": tests_per_frame * F3 / ms * 1e3,
        "hbm_algorithmic_gbs": b_alg / ms / 1e6, "hbm_frac": b_alg / ms / 1e6 / PEAK_HBM,
        "path": "global-memory cell path (5 launches per frame), dense 40 GB atom table",
        "extrapolated_1M_frames_s": 1e6 / (F3 / ms * 1e3)}
    del atom_t, res_t
    torch.cuda.empty_cache()

    # ---- S4: min-dist on S3 ---------------------------------------------------------------
    F4 = 16 if quick else 100
    xyz4, box4 = xyz[:F4].contiguous(), box[:F4].contiguous()
    ligand_q = np.arange(0, 30)
    solvent_h = np.arange(2700, system.n_atoms)
    ms, (dmin, darg) = timed(lambda: engine.min_dist(xyz4, box4, ligand_q, solvent_h, False))
    out["S4_min_dist_ligand30_x_solvent"] = {
        "frames": F4, "pairs_per_frame": int(len(ligand_q) * len(solvent_h)), "ms": ms,
        "frames_per_s": F4 / ms * 1e3, "pairs_per_s": len(ligand_q) * len(solvent_h) * F4 / ms * 1e3,
        "arithmetic": "dist_mic_triclinic (27 image search)"}
    prot_q = np.arange(0, 2700)
    F4b = 2 if quick else 8
    ms, _ = timed(lambda: engine.min_dist(xyz4[:F4b], box4[:F4b], prot_q, solvent_h, False), repeat=1)
    out["S4_min_dist_protein2700_x_solvent"] = {
        "frames": F4b, "pairs_per_frame": int(len(prot_q) * len(solvent_h)), "ms": ms,
        "frames_per_s": F4b / ms * 1e3, "pairs_per_s": len(prot_q) * len(solvent_h) * F4b / ms * 1e3,
        "note": "the reference would materialise 1 GB of float32 per frame"}
    ms, _ = timed(lambda: engine.nearest(xyz4[0], box4[0].reshape(9), 0.45, False, 1, atom_res=res), repeat=2)
    out["S4_nearest_atoms_100k"] = {"ms": ms, "atoms": system.n_atoms, "method": "tiled brute force, 1e10 pair tests",
                                    "pair_tests_per_s": system.n_atoms ** 2 / ms * 1e3}

    # ---- chunked runner on S1 (reference partition default_slices(F, 8)) ------------------
    s1 = synth.kras_like()
    F1 = 2000 if quick else 10000
    engine.set_system(*s1.topology.atom_residue_chain(), np.full(s1.n_atoms, 3, np.uint8), 0.45, 2)
    xyz1, box1 = device_frames(engine, s1, 0, F1)
    traj1 = DevTraj(s1.topology, xyz1, box1)
    everyone = np.arange(s1.n_atoms)
    cm.ChunkedContactFrequency(traj1, query=everyone, haystack=everyone, n_blocks=8)
    t0 = time.perf_counter()
    chunked = cm.ChunkedContactFrequency(traj1, query=everyone, haystack=everyone, n_blocks=8)
    t1 = time.perf_counter()
    serial = cm.ContactFrequency(traj1, query=everyone, haystack=everyone)
    t2 = time.perf_counter()
    same = (chunked._table_of("atom").value.sum() == serial._table_of("atom").value.sum()
            and len(chunked._table_of("atom")) == len(serial._table_of("atom")))
    out["S1_chunked_runner_8_blocks"] = {"frames": F1, "chunked_s": t1 - t0, "serial_s": t2 - t1,
                                         "chunked_frames_per_s": F1 / (t1 - t0), "equal_to_serial": bool(same)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
