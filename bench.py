#!/usr/bin/env python
"""Benchmark of the contact-map hot path on B200 (BASELINE.json metric / config).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU path (oracle port)

Workload (config.workload): "S1" = BASELINE.json configs[1], synthetic KRAS-sized protein,
2700 heavy atoms, 10 000 frames, orthorhombic PBC, query = haystack = all atoms, cutoff
0.45 nm, n_neighbors_ignored = 2.  A "step" is one full ContactFrequency pass over the 10 000
frames: zero the count tables, bin + screen + accumulate every frame, (N > 1: one
NCCL reduce of the tables to rank 0), sparse export of both tables.  Multi-GPU runs shard by frame with
fixed per-GPU work (weak scaling): every rank owns 10 000 different frames.

`value` is measured with CUDA events, inputs resident in HBM (generated on device by the
bit-reproducible generator).  `e2e` goes through the public API
(`ChunkedContactFrequency(trajectory)`) with the coordinates in PINNED HOST memory: host->device
copies, kernels, reduce, export and the device->host read of the sparse result are all inside
the timed region.  `roofline` is the HBM view of the dominant kernel: algorithmic bytes of one
launch over that kernel's own duration, taken from CUDA events the library records around every
`vl_frames` launch on the stream it launches on (`kernel_ms`; `accumulate_ms` is the whole
`cmb_accumulate`, list build included); `roofline.fp32_pipe` is SURVEY 8d's work-equivalent FP32
view.  `legs` (N = 1) are short runs of the other BASELINE.json configs; `--workload s3` makes
configs[3] (100 000 atoms) the timed workload.
"""
import os
import sys

if "--impl" in sys.argv and "reference" in sys.argv:
    # the reference arm runs one oracle process per core: the OpenMP team size has to be fixed
    # BEFORE anything loads libgomp (numpy / torch / the oracle library read it at load time)
    os.environ["OMP_NUM_THREADS"] = "1"

import argparse
import json
import statistics
import subprocess
import threading
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "frames/sec for ContactFrequency @0.45 nm"
N_FRAMES = 10000
CUTOFF = 0.45
N_IGNORED = 2


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="s1", choices=["s1", "s3"],
                    help="s1: BASELINE.json configs[1] (the headline); s3: configs[3], 100 000-atom triclinic system")
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU (default 10000 for s1, 4096 for s3)")
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames of the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-legs", action="store_true", help="skip the S2 / S3 / S4 legs of the default run")
    ap.add_argument("--no-pair-list", action="store_true", help="cell kernels only (no frame-batch pair list)")
    args = ap.parse_args()
    if args.frames <= 0:
        args.frames = N_FRAMES if args.workload == "s1" else 4096
    return args


def workload_config(args, world):
    if args.workload == "s3":
        return {
            "workload": "S3: synthetic solvated 100 000-atom triclinic-box system (32 704 residues), all-atom "
                        "contacts, %d frames per GPU generated on device, cutoff 0.45 nm, n_neighbors_ignored 2 "
                        "(BASELINE.json configs[3])" % args.frames,
            "frames_per_gpu": args.frames, "n_atoms": 100000, "cutoff_nm": CUTOFF, "n_neighbors_ignored": N_IGNORED,
            "sharding": "frames (rank r owns frames [r*F, (r+1)*F)); the ranks' sparse count lists are all-gathered "
                        "and merged by key (the dense 40 GB tables are never moved)" if world > 1 else "single GPU",
            "l2": "inputs larger than L2: %.0f MB of coordinates per pass vs 126 MB L2" % (args.frames * 1.2),
        }
    return {
        "workload": "S1: synthetic KRAS-sized protein, 2700 heavy atoms (270 residues x 10), "
                    "%d frames per GPU, orthorhombic PBC 7.0x7.5x8.0 nm, query=haystack=all, "
                    "cutoff 0.45 nm, n_neighbors_ignored 2 (BASELINE.json configs[1])" % args.frames,
        "frames_per_gpu": args.frames,
        "n_atoms": 2700,
        "cutoff_nm": CUTOFF,
        "n_neighbors_ignored": N_IGNORED,
        "sharding": "frames (rank r owns frames [r*F, (r+1)*F)), one NCCL reduce of the count tables to rank 0"
                    if world > 1 else "single GPU",
        "l2": "inputs larger than L2: %.0f MB of coordinates per pass vs 126 MB L2" % (args.frames * 2700 * 12 / 1e6),
    }


# -------------------------------------------------------------------------------------
# clocks
# -------------------------------------------------------------------------------------
class ClockSampler(object):
    """SM clock / throttle-reason samples DURING the timed region.

    NVML (nvidia_ml_py) is polled from a thread every few milliseconds, so that even a timed
    region of some tens of milliseconds gets samples; `nvidia-smi -lms` is the fallback."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.samples = []          # (sm_mhz, reasons bitmask)
        self.lines = []
        self.proc = None
        self.thread = None
        self.stop_flag = False
        self.nvml = None
        self.handle = None
        self.sm_max = None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = gpu_index
            if visible:
                ids = [v.strip() for v in visible.split(",") if v.strip()]
                if gpu_index < len(ids) and ids[gpu_index].isdigit():
                    phys = int(ids[gpu_index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def start(self):
        self.stop_flag = False
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _poll_nvml(self):
        nv = self.nvml
        while not self.stop_flag:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                try:
                    reasons = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    reasons = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.samples.append((mhz, reasons))
            except Exception:
                pass
            time.sleep(0.003)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        self.stop_flag = True
        if self.nvml is not None:
            self.thread.join(timeout=1.0)
            nv = self.nvml
            names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            sm = [m for m, _ in self.samples]
            reasons = sorted(name for name, bit in names.items() if any(r & bit for _, r in self.samples))
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.sm_max,
                    "samples": len(sm), "reasons": reasons, "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[5:9]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi"}


# -------------------------------------------------------------------------------------
# CPU legs (the oracle is the reference's CPU path restated; this is the one place bench.py
# may execute oracle/)
# -------------------------------------------------------------------------------------
def cpu_reference_rate(system, n_frames, n_workers):
    """frames/s of the reference CPU path (oracle port) on `n_frames` frames of the workload.
    n_workers == 1: the serial ContactFrequency loop; > 1: the Dask-equivalent runner
    (default_slices -> map over a process pool -> Counter reduce)."""
    from oracle import refpath
    top = system.topology
    xyz = system.frames(0, n_frames)
    box = np.broadcast_to(system.box, (n_frames, 3, 3)).copy()
    everyone = list(range(top.n_atoms))
    args = (xyz, box, top.atom_residue, top.residue_chain, everyone, everyone, CUTOFF, N_IGNORED)
    if n_workers == 1:
        t0 = time.perf_counter()
        atoms, residues = refpath.contact_frequency(*args)
        dt = time.perf_counter() - t0
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(n_workers, initializer=_worker_init) as pool:
            # spin the workers up outside the timed region and record the OpenMP team size each
            # of them really runs the oracle with
            observed = pool.map(_worker_threads, range(n_workers))
            cpu_reference_rate.worker_threads = max(observed)
            t0 = time.perf_counter()
            atoms, residues = refpath.chunked_contact_frequency(*args, n_workers=n_workers, pool=pool)
            dt = time.perf_counter() - t0
    return n_frames / dt, dt, len(atoms), len(residues)


def _worker_init():
    """One OpenMP thread per worker process (one worker per core): omp_set_num_threads through
    the oracle library, which works whatever OMP_NUM_THREADS said when libgomp was loaded."""
    from oracle import clib
    clib.set_threads(1)


def _worker_threads(_):
    from oracle import clib
    return clib.max_threads()


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path (oracle port; MDTraj and
    dask are not installable here) with all host cores, on a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from contact_map_b200 import synth
    from oracle import clib
    clib.lib()                       # build / load once in the parent, before the workers fork
    system = synth.solvated_triclinic(n_atoms=100000) if args.workload == "s3" else synth.kras_like()
    cores = len(os.sched_getaffinity(0))
    workers = max(1, cores)
    frames = args.cpu_frames or (workers if args.workload == "s3" else max(workers * 16, 64))
    rates = []
    for _ in range(args.warmup):
        cpu_reference_rate(system, workers if args.workload == "s3" else max(workers, 8), workers)
    t_total = 0.0
    for _ in range(args.steps):
        rate, dt, _, _ = cpu_reference_rate(system, frames, workers)
        rates.append(rate)
        t_total += dt
    value = frames * args.steps / t_total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {
            "value": value, "unit": "frames/s", "cores": workers, "kind": "port",
            "sample": "%d frames of the workload per step; reference Python set/Counter algorithm restated "
                      "(oracle/refpath.py) over the C restatement of MDTraj's neighbour list, run as the "
                      "Dask-equivalent runner: default_slices -> %d worker processes -> Counter reduce "
                      "(oracle C lib OpenMP threads per worker, observed inside the workers: %d)"
                      % (frames, workers, getattr(cpu_reference_rate, "worker_threads", 0)),
        },
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "oracle_threads_per_worker": getattr(cpu_reference_rate, "worker_threads", None),
        "oracle_threads_parent": clib.max_threads(),
    }
    print(json.dumps(line))
    return 0


# -------------------------------------------------------------------------------------
# GPU arm
# -------------------------------------------------------------------------------------
FP32_OPS = {"ortho": 21.0, "triclinic": 27.0}     # 18 FP32 + 3 XU / 24 FP32 + 3 XU lane-ops per candidate


def device_frames(engine, system, frame0, n_frames):
    import torch
    dev = engine.device
    xyz = engine.synth_frames(torch.as_tensor(system.base, device=dev),
                              torch.as_tensor(system.groups.astype(np.int32), device=dev),
                              torch.as_tensor(system.box.reshape(9), device=dev), system.seed,
                              system.amplitude, True, frame0, n_frames)
    box = torch.as_tensor(system.box, device=dev).reshape(1, 3, 3).expand(n_frames, 3, 3).contiguous()
    return xyz, box


def roles_of(system):
    n = system.n_atoms
    role = np.zeros(n, dtype=np.uint8)
    role[np.arange(n) if system.query is None else system.query] |= 1
    role[np.arange(n) if system.haystack is None else system.haystack] |= 2
    return role


def pair_list_stats(engine):
    return {k[3:]: engine.get_stat(k) for k in ("vl_used", "vl_valid", "vl_tripped", "vl_entries", "vl_tiles",
                                                "vl_rows", "vl_slots", "vl_segs")}


def canonical_pairs(engine, xyz, box, atom_t, res_t):
    """Candidate pairs of one pass counted by the cell kernels (separate, untimed launches with
    the counter switched on): canonical = cutoff-sized cells, 27-cell half shell (SURVEY 8d's
    P_cand; task pairing off)."""
    engine.set_option("count_tests", 1)
    engine.set_option("pair_tasks", 0)
    atom_t.zero_()
    res_t.zero_()
    engine.accumulate(xyz, box, atom_t, res_t)
    n = engine.get_stat("pair_tests")
    engine.set_option("pair_tasks", 1)
    engine.set_option("count_tests", 0)
    return n


def h2d_probe(device, world, barrier):
    """Pinned host -> device copy rate of this rank, alone and with every rank copying at once
    (GB/s): the ceiling of the end-to-end number."""
    import torch
    import torch.distributed as dist
    nbytes = 256 << 20
    h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(nbytes, dtype=torch.uint8, device=device)
    out = {}
    for name in ("alone", "all_ranks"):
        rates = []
        for rank_turn in (range(world) if name == "alone" else [None]):
            barrier()
            if rank_turn is None or rank_turn == int(os.environ.get("RANK", "0")):
                d.copy_(h, non_blocking=True)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(4):
                    d.copy_(h, non_blocking=True)
                e1.record()
                torch.cuda.synchronize()
                rates.append(4 * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9)
        rate = min(rates)
        if world > 1:
            t = torch.tensor([rate], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            rate = float(t.item())
        out[name] = rate
    return {"h2d_peak_gbs": out["alone"], "h2d_peak_gbs_all_ranks_concurrent": out["all_ranks"],
            "h2d_probe": "4 x 256 MiB pinned->device copies per rank, CUDA events; 'alone' = slowest rank copying "
                         "by itself, 'concurrent' = slowest rank while all %d ranks copy" % world}


def timed_ms(fn, repeat=3):
    import torch
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


def extra_legs(engine, fp32_peak, budget_s=150.0):
    """Short, untimed-by-the-headline legs for the other BASELINE.json configs (N = 1 only):
    S2 ContactTrajectory (+ ContactDifference), S3 ContactFrequency on 100 000 atoms, S4
    MinimumDistanceCounter / NearestAtoms.  Device work timed with CUDA events, inputs resident."""
    import torch
    import contact_map_b200 as cm
    from contact_map_b200 import synth
    t_start = time.perf_counter()
    legs = {}
    # ---- S2: ligand and partner chain against the receptor --------------------------------
    for ligand in (True, False):
        name = "s2_ligand" if ligand else "s2_partner"
        system = synth.binding_system(ligand=ligand)
        top = system.topology
        res, chain = top.atom_residue_chain()
        engine.set_system(res, chain, roles_of(system), CUTOFF, N_IGNORED)
        F = 20000
        xyz, box = device_frames(engine, system, 0, F)
        engine.frame_contacts(xyz, box, frame0=0)               # sizes the key buffers
        ms, (ak, rk) = timed_ms(lambda: engine.frame_contacts(xyz, box, frame0=0), repeat=3)
        legs[name + "_contact_trajectory"] = {
            "frames": F, "n_atoms": system.n_atoms, "frames_per_s": F / ms * 1e3, "atom_keys": int(len(ak)),
            "residue_keys": int(len(rk)), "includes": "key emission kernel, radix sort, D2H of the sorted keys"}
        # public API from HOST memory: ContactTrajectory + ContactDifference(first half, second half)
        h_xyz = torch.empty((F, system.n_atoms, 3), dtype=torch.float32, pin_memory=True)
        h_xyz.copy_(xyz)
        h_box = torch.empty((F, 3, 3), dtype=torch.float32, pin_memory=True)
        h_box.copy_(box)
        torch.cuda.synchronize()
        traj = cm.Trajectory(h_xyz.numpy(), top, h_box.numpy())
        kw = dict(query=system.query, haystack=system.haystack, cutoff=CUTOFF, n_neighbors_ignored=N_IGNORED)
        cm.ContactTrajectory(traj, **kw)
        t0 = time.perf_counter()
        ct = cm.ContactTrajectory(traj, **kw)
        t1 = time.perf_counter()
        diff = cm.ContactFrequency(traj[:F // 2], **kw) - cm.ContactFrequency(traj[F // 2:], **kw)
        n_diff = len(diff.atom_contacts.counter)
        t2 = time.perf_counter()
        legs[name + "_api_host"] = {"frames": F, "ContactTrajectory_frames_per_s": F / (t1 - t0),
                                    "graph_captures": engine.get_stat("vl_graph_captures"),
                                    "graph_capture_ms": engine.get_stat("vl_graph_capture_us") * 1e-3,
                                    "two_ContactFrequency_plus_ContactDifference_s": t2 - t1,
                                    "difference_pairs": n_diff, "frames_in_trajectory": len(ct)}
        del xyz, box, h_xyz, h_box, traj, ct
        if time.perf_counter() - t_start > budget_s:
            return legs
    # ---- S3: 100 000-atom triclinic system ----------------------------------------------------
    system = synth.solvated_triclinic(n_atoms=100000)
    top = system.topology
    res, chain = top.atom_residue_chain()
    engine.set_system(res, chain, roles_of(system), CUTOFF, N_IGNORED)
    F3 = 512
    xyz, box = device_frames(engine, system, 0, F3)
    atom_t, res_t = engine.new_tables()
    engine.accumulate(xyz[:8], box[:8], atom_t, res_t)           # page the tables in
    ms, _ = timed_ms(lambda: engine.accumulate(xyz, box, atom_t, res_t), repeat=3)
    stats = pair_list_stats(engine)
    engine.set_option("verlet", 0)
    ms_cells, _ = timed_ms(lambda: engine.accumulate(xyz[:64], box[:64], atom_t, res_t), repeat=1)
    engine.set_option("verlet", 1)
    cand = canonical_pairs(engine, xyz[:4].contiguous(), box[:4].contiguous(), atom_t, res_t) / 4.0
    b_alg = F3 * (12 * system.n_atoms + 36)
    legs["s3_contact_frequency"] = {
        "frames": F3, "n_atoms": system.n_atoms, "n_residues": top.n_residues, "frames_per_s": F3 / ms * 1e3,
        "ms": ms, "includes": "pair-list build + guard pass + frame kernel + fall back", "pair_list": stats,
        "cell_kernels_only_frames_per_s": 64 / ms_cells * 1e3,
        "candidate_pairs_per_frame": cand, "hbm_algorithmic_gbs": b_alg / ms / 1e6,
        "fp32_frac": (F3 / ms * 1e3) * cand * FP32_OPS["triclinic"] / fp32_peak,
        "fp32_definition": "frames/s x canonical candidate pairs x 27 lane-ops (triclinic, SURVEY 8d) / FP32 pipe peak"}
    del atom_t, res_t
    torch.cuda.empty_cache()
    # ---- S4: min-dist / nearest on the 100k system --------------------------------------------
    F4 = 64
    xyz4, box4 = xyz[:F4].contiguous(), box[:F4].contiguous()
    lig, solvent, protein = np.arange(0, 30), np.arange(2700, system.n_atoms), np.arange(0, 2700)
    ms, _ = timed_ms(lambda: engine.min_dist(xyz4, box4, lig, solvent, False), repeat=2)
    legs["s4_min_dist_ligand30_x_solvent"] = {"frames": F4, "pairs_per_frame": int(len(lig) * len(solvent)),
                                              "frames_per_s": F4 / ms * 1e3}
    ms, _ = timed_ms(lambda: engine.min_dist(xyz4[:8], box4[:8], protein, solvent, False), repeat=1)
    legs["s4_min_dist_protein2700_x_solvent"] = {"frames": 8, "pairs_per_frame": int(len(protein) * len(solvent)),
                                                 "frames_per_s": 8 / ms * 1e3}
    ms, _ = timed_ms(lambda: engine.nearest(xyz4[0], box4[0].reshape(9), CUTOFF, False, 1, atom_res=res), repeat=2)
    legs["s4_nearest_atoms_100k"] = {"ms": ms, "atoms": system.n_atoms}
    return legs


def run_ours(args):
    import torch
    import torch.distributed as dist
    warnings.simplefilter("ignore", PendingDeprecationWarning)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout at communicator creation; keep stdout for the
        # single JSON line by pointing fd 1 at stderr until the first collective has run
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            warm = torch.zeros(1, device=device)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    import contact_map_b200 as cm
    from contact_map_b200 import parallel, synth
    from contact_map_b200.engine import get_engine

    s3 = args.workload == "s3"
    system = synth.solvated_triclinic(n_atoms=100000) if s3 else synth.kras_like()
    box_kind = "triclinic" if s3 else "ortho"
    top = system.topology
    n = top.n_atoms
    F = args.frames
    frame0 = rank * F
    engine = get_engine(local_rank)
    atom_res, atom_chain = top.atom_residue_chain()
    role = np.full(n, 3, dtype=np.uint8)          # every atom is query and haystack
    engine.set_system(atom_res, atom_chain, role, CUTOFF, N_IGNORED)
    engine.set_option("count_tests", 0)
    if args.no_pair_list:
        engine.set_option("verlet", 0)

    # inputs resident in HBM, generated on device (bit-identical to synth.make_frames)
    d_xyz, d_box = device_frames(engine, system, frame0, F)
    atom_t, res_t = engine.new_tables()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    kern_events = []

    def step(record):
        atom_t.zero_()
        res_t.zero_()
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        engine.accumulate(d_xyz, d_box, atom_t, res_t)
        if record:
            e1.record()
            kern_events.append((e0, e1))
        # the reference's reduce hands the result to the client only (frequency_task.py:125-141):
        # the count tables are summed onto rank 0, which alone holds the sparse result
        return parallel.reduce_export(engine, [atom_t, res_t], dst=0)

    for _ in range(max(args.warmup, 3)):      # timing rule: at least 3 warm-up steps
        out = step(False)
    n_atom_pairs, n_res_pairs = (len(out[0][0]), len(out[1][0])) if rank == 0 else (0, 0)

    launches0 = engine.get_stat("kernel_launches")
    # CUDA events around every launch of the pair-list frame kernel, recorded by the library on the
    # stream it launches on (torch.cuda.Event only sees torch's current stream)
    engine.set_option("time_kernels", 1)
    engine.get_stat("vl_frames_ns")
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(args.steps):
        step(True)
    stop.record()
    barrier()
    clocks = sampler.stop()
    elapsed_ms = start.elapsed_time(stop)
    launches = engine.get_stat("kernel_launches") - launches0
    kernel_ms = statistics.mean(e0.elapsed_time(e1) for e0, e1 in kern_events)
    vl_timed = engine.get_stat("vl_frames_timed")
    vl_ms = engine.get_stat("vl_frames_ns") * 1e-6 / vl_timed if vl_timed else None
    engine.set_option("time_kernels", 0)
    list_stats = pair_list_stats(engine)
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    value = world * F * args.steps / (elapsed_ms / 1e3)

    # canonical candidate pairs of one pass (SURVEY 8d's P_cand), counted on a sample of frames
    sample = min(F, 64 if s3 else 2000)
    pair_tests = canonical_pairs(engine, d_xyz[:sample].contiguous(), d_box[:sample].contiguous(), atom_t, res_t) \
        * (F / float(sample))

    # ---- end to end through the public API, host buffers -------------------------------
    e2e = None
    if not args.no_e2e:
        h_xyz = torch.empty((F, n, 3), dtype=torch.float32, pin_memory=True)
        h_xyz.copy_(d_xyz)
        h_box = torch.empty((F, 3, 3), dtype=torch.float32, pin_memory=True)
        h_box.copy_(d_box)
        torch.cuda.synchronize()
        traj = cm.Trajectory(h_xyz.numpy(), top, h_box.numpy())   # views of the pinned buffers
        assert traj.xyz.ctypes.data == h_xyz.data_ptr()
        everyone = np.arange(n)
        del atom_t, res_t                                           # (S3: 44 GB; the API allocates its own)
        torch.cuda.empty_cache()

        def e2e_step():
            return cm.ChunkedContactFrequency(traj, query=everyone, haystack=everyone, cutoff=CUTOFF,
                                              n_neighbors_ignored=N_IGNORED, presharded=True, result_rank=0)

        for _ in range(2):
            m = e2e_step()
        e2e_steps = args.steps if not s3 else min(args.steps, 3)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            m = e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        na, nr = len(m._table_of("atom")), len(m._table_of("residue"))
        e2e = {"value": world * F * e2e_steps / dt, "unit": "frames/s",
               "h2d_bytes_per_step": int(F * (n * 12 + 36)),
               "d2h_bytes_per_step": int((na + nr) * 12 + 16),
               "ms_per_step": 1e3 * dt / e2e_steps, "steps": e2e_steps,
               "api": "ChunkedContactFrequency(trajectory with pinned host xyz, presharded=True) -> "
                      "cmb_accumulate_host (chunked H2D overlapped with kernels) -> reduce to rank 0 -> sparse export -> D2H"}
        e2e.update(h2d_probe(device, world, barrier))
        # the copy is the floor of the end-to-end step: fraction of the measured H2D ceiling reached
        e2e["h2d_gbs_achieved_per_gpu"] = e2e["h2d_bytes_per_step"] / (e2e["ms_per_step"] * 1e-3) / 1e9
        e2e["frac_of_h2d_peak_concurrent"] = e2e["h2d_gbs_achieved_per_gpu"] / e2e["h2d_peak_gbs_all_ranks_concurrent"]
        del traj, h_xyz, h_box, m

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel -------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    alg_bytes = F * (12 * n + 36)
    achieved = alg_bytes / (kernel_ms / 1e3) / 1e9
    traffic = None
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    dominant = "vl_frames" if list_stats["used"] and list_stats["valid"] else ("kl_pairs" if s3 else "k_frames_small")
    if os.path.exists(traffic_path):
        per_frame = json.load(open(traffic_path)).get("%s_%s_dram_bytes_per_frame" % (dominant, args.workload))
        traffic = None if per_frame is None else per_frame * F    # ncu capture, scaled to this launch
    sm_count = engine.get_stat("sm_count")
    sm_mhz = clocks.get("sm_mhz") or 1965.0
    # FP32 view: canonical candidate pairs x the lane-ops of the reference arithmetic
    # (orthorhombic 18 FP32 + 3 XU = 21, triclinic 24 + 3 = 27) against the FP32 pipe peak.  The
    # kernels screen with FMA arithmetic and run the reference arithmetic only on the pairs the
    # screening cannot decide, and the pair list skips most of the canonical candidates altogether,
    # so this is work-equivalent throughput, not pipe utilisation.
    ops = FP32_OPS[box_kind]
    fp32_lane_ops = pair_tests * ops / (kernel_ms / 1e3)
    fp32_peak = sm_count * 128 * sm_mhz * 1e6
    if dominant == "vl_frames" and vl_ms:
        # the dominant kernel alone: one vl_frames launch per step reads every frame once
        accumulate_ms, kernel_ms = kernel_ms, vl_ms
        achieved = alg_bytes / (kernel_ms / 1e3) / 1e9
        covers = "vl_frames alone (CUDA events recorded by the library around each launch, on its stream)"
    else:
        accumulate_ms = kernel_ms
        covers = "cmb_accumulate: frame kernels of the cell path"
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "kernel": dominant, "kernel_ms": kernel_ms,
        "kernel_ms_covers": covers,
        "accumulate_ms": accumulate_ms,
        "accumulate_ms_covers": "cmb_accumulate: pair-list build (~30 small launches), guard, frame kernel, fall back",
        "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
        "note": "HBM is the designated roofline (SURVEY 8d) but the path is instruction / shared-memory bound: "
                "see fp32_pipe (SURVEY 8d's lower ceiling)",
        "fp32_pipe": {"candidate_pairs_per_launch": pair_tests, "reference_lane_ops_per_pair": ops,
                      "achieved_lane_ops_per_s": fp32_lane_ops,
                      "peak_lane_ops_per_s": fp32_peak, "frac": fp32_lane_ops / fp32_peak,
                      "time_basis": "accumulate_ms (list build included)",
                      "executed": (None if not (dominant == "vl_frames" and list_stats.get("entries")) else {
                          "listed_pairs_per_launch": float(list_stats["entries"]) * F,
                          "listed_pairs_per_s": float(list_stats["entries"]) * F / (kernel_ms / 1e3),
                          "note": "what the frame kernel actually walks: the skin-padded pair list (one 6-flop screen "
                                  "per listed pair and frame), against candidate_pairs_per_launch canonical candidates; "
                                  "a work-equivalent frac above 1 means the list skips more work than the pipe could do"}),
                      "definition": "frames/s x canonical candidate pairs (cutoff-sized cells, 27-cell half "
                                    "shell) x %d lane-ops (SURVEY 8d, %s) of the reference arithmetic / FP32 pipe "
                                    "peak; work-equivalent throughput, not pipe utilisation" % (ops, box_kind),
                      "peak_definition": "%d SMs x 128 FP32 lanes x %.0f MHz (median SM clock under load)"
                                         % (sm_count, sm_mhz)},
    }

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        frames_cpu = args.cpu_frames or (1 if s3 else 512)
        rate, dt, _, _ = cpu_reference_rate(system, frames_cpu, 1)
        cpu_baseline = {"value": rate, "unit": "frames/s", "cores": 1, "kind": "port",
                        "sample": "first %d frames of %s (%.1f s): reference Python set/Counter loop restated "
                                  "(oracle/refpath.py) over the C restatement of MDTraj's neighbour list "
                                  "(oracle/cm_oracle.c); serial ContactFrequency, as the reference runs it"
                                  % (frames_cpu, args.workload.upper(), dt)}

    legs = None
    if world == 1 and not s3 and not args.no_legs:
        del d_xyz, d_box
        torch.cuda.empty_cache()
        legs = extra_legs(engine, fp32_peak)

    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "cpu_baseline": cpu_baseline,
        "pair_list": list_stats,
        "atom_pairs_screened_per_s": {"logical_nQ_x_nH": value * n * n,
                                      "candidate_pairs": world * pair_tests * args.steps / (elapsed_ms / 1e3)},
        "result": {"distinct_atom_pairs": n_atom_pairs, "distinct_residue_pairs": n_res_pairs},
        "legs": legs,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
