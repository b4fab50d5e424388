"""torchrun --nproc-per-node N tools/check_multirank_s3.py [frames]: BASELINE configs[3] on N ranks.
The 100 000-atom triclinic system, a prefix of `frames` (default 64) frames sharded over the ranks
by the reference partition (frequency_task.default_slices), per-rank dense 40 GB tables, the
sparse reduce (parallel.reduce_export: every rank's non-zero entries are all-gathered and merged by
key) -- against the tables one rank computes alone from all the frames.  Both frame paths: cell
kernels (verlet 0) and pair list (verlet 2)."""
import json, os, sys, time, warnings
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from contact_map_b200 import parallel, synth
from contact_map_b200.engine import get_engine
F = int(sys.argv[1]) if len(sys.argv) > 1 else 64
system = synth.solvated_triclinic(n_atoms=100000)
engine = get_engine(local); dev = engine.device
res, chain = system.topology.atom_residue_chain()
engine.set_system(res, chain, np.full(system.n_atoms, 3, np.uint8), 0.45, 2)
xyz = engine.synth_frames(torch.as_tensor(system.base, device=dev), torch.as_tensor(system.groups.astype(np.int32), device=dev),
                          torch.as_tensor(system.box.reshape(9), device=dev), system.seed, system.amplitude, True, 0, F)
box = torch.as_tensor(system.box, device=dev).reshape(1, 3, 3).expand(F, 3, 3).contiguous()
atom_t, res_t = engine.new_tables()
ok = True
out = {"world": world, "frames": F}
for mode in (0, 2):
    engine.set_option("verlet", mode)
    atom_t.zero_(); res_t.zero_()
    engine.accumulate(xyz, box, atom_t, res_t)
    serial = engine.export_many([atom_t, res_t])
    atom_t.zero_(); res_t.zero_()
    mine = 0
    for s in parallel.rank_slices(F, n_blocks=max(world, 4)):
        engine.accumulate(xyz[s].contiguous(), box[s].contiguous(), atom_t, res_t)
        mine += s.stop - s.start
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    merged = parallel.reduce_export(engine, [atom_t, res_t], dst=None)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    same = all(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) for a, b in zip(serial, merged))
    total = parallel.reduce_frame_count(mine, dev)
    ok = ok and same and total == F
    out["verlet%d" % mode] = {"equal": bool(same), "frames_total": int(total), "frames_this_rank": mine,
                              "atom_pairs": int(len(merged[0][0])), "res_pairs": int(len(merged[1][0])),
                              "reduce_export_ms": dt * 1e3, "list_used": int(engine.get_stat("vl_used"))}
print("rank %d %s" % (rank, json.dumps(out)), flush=True)
if world > 1:
    dist.destroy_process_group()
print("CHECK_MULTIRANK_S3", "OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
