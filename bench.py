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
the timed region.
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
    ap.add_argument("--frames", type=int, default=N_FRAMES, help="frames per GPU (default 10000)")
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames of the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload_config(args, world):
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
    system = synth.kras_like()
    cores = len(os.sched_getaffinity(0))
    workers = max(1, cores)
    frames = args.cpu_frames or max(workers * 16, 64)
    rates = []
    for _ in range(args.warmup):
        cpu_reference_rate(system, max(workers, 8), workers)
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
            "sample": "%d frames of S1 per step; reference Python set/Counter algorithm restated "
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

    system = synth.kras_like()
    top = system.topology
    n = top.n_atoms
    F = args.frames
    frame0 = rank * F
    engine = get_engine(local_rank)
    atom_res, atom_chain = top.atom_residue_chain()
    role = np.full(n, 3, dtype=np.uint8)          # every atom is query and haystack
    engine.set_system(atom_res, atom_chain, role, CUTOFF, N_IGNORED)
    engine.set_option("count_tests", 0)

    # inputs resident in HBM, generated on device (bit-identical to synth.make_frames)
    d_base = torch.as_tensor(system.base, device=device)
    d_groups = torch.as_tensor(system.groups.astype(np.int32), device=device)
    d_box9 = torch.as_tensor(system.box.reshape(9), device=device)
    d_xyz = engine.synth_frames(d_base, d_groups, d_box9, system.seed, system.amplitude, True, frame0, F)
    d_box = d_box9.reshape(1, 3, 3).expand(F, 3, 3).contiguous()
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
        # one NCCL reduce to rank 0, which alone exports it
        parallel.reduce_tables([atom_t, res_t], dst=0)
        return engine.export_many([atom_t, res_t]) if rank == 0 else None

    for _ in range(max(args.warmup, 3)):      # timing rule: at least 3 warm-up steps
        out = step(False)
    n_atom_pairs, n_res_pairs = (len(out[0][0]), len(out[1][0])) if rank == 0 else (0, 0)

    launches0 = engine.get_stat("kernel_launches")
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
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    value = world * F * args.steps / (elapsed_ms / 1e3)

    # candidate pairs of one pass (separate, untimed launches with the counter switched on):
    # canonical = cutoff-sized cells, 27-cell half shell (SURVEY 8d's P_cand; task pairing off);
    # screened = what the production configuration actually screens (task pairing on adds tests
    # between cells two columns apart)
    engine.set_option("count_tests", 1)
    counts = {}
    for key, pairing in (("canonical", 0), ("screened", 1)):
        engine.set_option("pair_tasks", pairing)
        atom_t.zero_()
        res_t.zero_()
        engine.accumulate(d_xyz, d_box, atom_t, res_t)
        counts[key] = engine.get_stat("pair_tests")
    engine.set_option("count_tests", 0)
    pair_tests = counts["canonical"]

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

        def e2e_step():
            m = cm.ChunkedContactFrequency(traj, query=everyone, haystack=everyone, cutoff=CUTOFF,
                                           n_neighbors_ignored=N_IGNORED, presharded=True, result_rank=0)
            return m

        for _ in range(2):
            m = e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            m = e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        na, nr = len(m._table_of("atom")), len(m._table_of("residue"))
        e2e = {"value": world * F * args.steps / dt, "unit": "frames/s",
               "h2d_bytes_per_step": int(F * (n * 12 + 36)),
               "d2h_bytes_per_step": int((na + nr) * 12 + 16),
               "ms_per_step": 1e3 * dt / args.steps,
               "api": "ChunkedContactFrequency(trajectory with pinned host xyz, presharded=True) -> "
                      "cmb_accumulate_host (chunked H2D overlapped with kernels) -> NCCL reduce to rank 0 -> sparse export -> D2H"}

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
    if os.path.exists(traffic_path):
        per_frame = json.load(open(traffic_path)).get("k_frames_small_dram_bytes_per_frame")
        traffic = None if per_frame is None else per_frame * F    # ncu capture, scaled to this launch
    sm_count = engine.get_stat("sm_count")
    sm_mhz = clocks.get("sm_mhz") or 1965.0
    # FP32 view: canonical candidate pairs x 21 lane-ops of the reference arithmetic
    # (orthorhombic: 18 FP32 + 3 XU per pair; triclinic would be 24 + 3), against the FP32 pipe peak.  The
    # kernel itself spends 10 issue slots per screened pair (3 FADD + 3 FFMA + 2 FADD + FMNMX + SHF)
    # and runs the reference arithmetic only on the pairs the screening cannot decide.
    fp32_lane_ops = pair_tests * 21.0 / (kernel_ms / 1e3)
    fp32_peak = sm_count * 128 * sm_mhz * 1e6
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "kernel": "k_frames_small", "kernel_ms": kernel_ms,
        "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
        "note": "HBM is the designated roofline (SURVEY 8d) but the kernel is instruction-issue bound: "
                "see fp32_pipe (SURVEY 8d's lower ceiling)",
        "fp32_pipe": {"candidate_pairs_per_launch": pair_tests,
                      "screened_pairs_per_launch": counts["screened"],
                      "reference_lane_ops_per_pair": 21, "kernel_issue_slots_per_screened_pair": 10,
                      "achieved_lane_ops_per_s": fp32_lane_ops,
                      "peak_lane_ops_per_s": fp32_peak, "frac": fp32_lane_ops / fp32_peak,
                      "definition": "frames/s x canonical candidate pairs (cutoff-sized cells, 27-cell half "
                                    "shell) x 21 lane-ops (18 FP32 + 3 XU, SURVEY 8d) of the reference arithmetic / FP32 pipe peak; the "
                                    "kernel screens with FMA arithmetic and adjudicates only ambiguous pairs "
                                    "with the reference arithmetic, so frac is work-equivalent throughput, "
                                    "not pipe utilisation",
                      "peak_definition": "%d SMs x 128 FP32 lanes x %.0f MHz (median SM clock under load)"
                                         % (sm_count, sm_mhz)},
    }

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        frames_cpu = args.cpu_frames or 512
        rate, dt, _, _ = cpu_reference_rate(system, frames_cpu, 1)
        cpu_baseline = {"value": rate, "unit": "frames/s", "cores": 1, "kind": "port",
                        "sample": "first %d frames of S1 (%.1f s): reference Python set/Counter loop restated "
                                  "(oracle/refpath.py) over the C restatement of MDTraj's neighbour list "
                                  "(oracle/cm_oracle.c); serial ContactFrequency, as the reference runs it"
                                  % (frames_cpu, dt)}

    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "cpu_baseline": cpu_baseline,
        "atom_pairs_screened_per_s": {"logical_nQ_x_nH": value * n * n,
                                      "candidate_pairs": world * pair_tests * args.steps / (elapsed_ms / 1e3),
                                      "screened_pairs": world * counts["screened"] * args.steps / (elapsed_ms / 1e3)},
        "result": {"distinct_atom_pairs": n_atom_pairs, "distinct_residue_pairs": n_res_pairs},
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
