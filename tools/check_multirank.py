"""torchrun --nproc-per-node 2 tools/check_multirank.py: the sharded runner on 2 ranks (NCCL), dense
and hashed tables, against the single-rank result of the same trajectory."""
import os, sys, warnings
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl")
import contact_map_b200 as cm
from contact_map_b200 import synth
from contact_map_b200.engine import ContactEngine, get_engine
rng = np.random.default_rng(5)
system = synth.random_system(rng, 600, n_chains=2, box_kind="triclinic", extent=2.4)
traj = system.trajectory(23)
engine = get_engine(local)
serial = cm.ContactFrequency(traj, cutoff=0.45, n_neighbors_ignored=2)
ok = True
for hashed in (False, True):
    if hashed:
        engine.set_option("force_path", 1)
        ContactEngine.DENSE_LIMIT_BYTES = 1024
    sharded = cm.ChunkedContactFrequency(traj, cutoff=0.45, n_neighbors_ignored=2, n_blocks=5)
    same = (sharded._atom_contacts == serial._atom_contacts and sharded._residue_contacts == serial._residue_contacts
            and sharded.n_frames == len(traj))
    print("rank %d hashed=%s tables=%s equal=%s" % (rank, hashed, sharded._table_kind, same))
    ok = ok and same
# reduce to rank 0 only: rank 0 holds the serial result, the other ranks an empty map
engine.set_option("force_path", -1)
ContactEngine.DENSE_LIMIT_BYTES = 1 << 36
rooted = cm.ChunkedContactFrequency(traj, cutoff=0.45, n_neighbors_ignored=2, n_blocks=5, result_rank=0)
if rank == 0:
    same = rooted._atom_contacts == serial._atom_contacts and rooted._residue_contacts == serial._residue_contacts
else:
    same = len(rooted._atom_contacts) == 0 and len(rooted._residue_contacts) == 0
print("rank %d result_rank=0 equal=%s" % (rank, same))
ok = ok and same
dist.destroy_process_group()
sys.exit(0 if ok else 1)
