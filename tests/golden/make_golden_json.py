"""Generate tests/golden/ref_json.json by running the UNMODIFIED reference (build container only):

    cd /root/reference && PYTHONDONTWRITEBYTECODE=1 \
        PYTHONPATH=/root/repo/oracle:/root/repo python /root/repo/tests/golden/make_golden_json.py

The fixture holds the reference's own ``to_json`` / ``to_dict`` output for ContactFrequency,
ContactTrajectory and ContactDifference on its bundled ``trajectory.pdb`` (cutoff 0.075, n = 0)
so that the drop-in's loaders are tested against files the reference wrote.  The script also
checks the opposite direction here, where the reference is importable: JSON written by
``contact_map_b200`` must load in the reference's ``from_json`` with identical counters.
"""
import json
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import mdtraj as md                      # noqa: E402  (the oracle stand-in)
import contact_map as ref                # noqa: E402  (the unmodified reference)
import contact_map_b200 as cm            # noqa: E402

warnings.simplefilter("ignore")
assert "oracle" in md.__file__ and ref.__file__.startswith("/root/reference")

traj = md.load("/root/reference/contact_map/tests/trajectory.pdb")
freq = ref.ContactFrequency(traj, cutoff=0.075, n_neighbors_ignored=0)
head = ref.ContactFrequency(traj[:2], cutoff=0.075, n_neighbors_ignored=0)
ctraj = ref.ContactTrajectory(traj, cutoff=0.075, n_neighbors_ignored=0)
diff = freq - head
out = {
    "frequency_json": freq.to_json(),
    "trajectory_dict": ctraj.to_dict(),
    "difference_dict": diff.to_dict(),
    "expected": {
        "n_frames": freq.n_frames,
        "atom_contacts": sorted([sorted(int(v) for v in k), v] for k, v in freq._atom_contacts.items()),
        "residue_contacts": sorted([sorted(int(v) for v in k), v] for k, v in freq._residue_contacts.items()),
        "difference_atom": sorted([sorted(int(v) for v in k), v] for k, v in diff.atom_contacts.counter.items()),
    },
}
with open(os.path.join(HERE, "ref_json.json"), "w") as f:
    json.dump(out, f)

# opposite direction: the drop-in's JSON in the reference's loaders
ours = cm.ContactFrequency.from_json(out["frequency_json"])
back = ref.ContactFrequency.from_json(ours.to_json())
assert back._atom_contacts == freq._atom_contacts and back._residue_contacts == freq._residue_contacts
assert back.n_frames == freq.n_frames and back.cutoff == freq.cutoff
assert set(back.query) == set(freq.query) and back.topology == freq.topology
assert set(json.loads(ours.to_json())) == set(json.loads(out["frequency_json"]))
ours_t = cm.ContactTrajectory.from_dict(out["trajectory_dict"])
back_t = ref.ContactTrajectory.from_dict(ours_t.to_dict())
assert [m._atom_contacts for m in back_t] == [m._atom_contacts for m in ctraj]
ours_d = cm.ContactDifference.from_dict(out["difference_dict"])
back_d = ref.ContactDifference.from_dict(ours_d.to_dict())
assert back_d.atom_contacts.counter == diff.atom_contacts.counter
print("wrote ref_json.json; the reference loads contact_map_b200's JSON with identical results")
