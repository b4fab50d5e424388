"""Access to the committed golden fixtures (tests/golden/*.npz; see make_golden.py)."""
import collections
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CASE_NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                    if os.path.basename(p).startswith(("pdb_c", "syn_")))


def load(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))


def counter(pairs, values):
    return collections.Counter({frozenset((int(a), int(b))): (v.item() if hasattr(v, "item") else v)
                                for (a, b), v in zip(pairs, values)})


def frame_sets(ptr, pairs):
    return [set(frozenset((int(a), int(b))) for a, b in pairs[ptr[f]:ptr[f + 1]])
            for f in range(len(ptr) - 1)]


def trajectory(case):
    """Product-side Trajectory built from a fixture's inputs."""
    from contact_map_b200.trajectory import Topology, Trajectory
    top = Topology(case["atom_residue"], case["residue_chain"],
                   atom_elements=[str(s) for s in case["atom_elements"]],
                   residue_names=[str(s) for s in case["residue_names"]])
    return Trajectory(case["xyz"], top, case.get("box"))


def call_kwargs(case):
    kwargs = {"cutoff": float(case["cutoff"]), "n_neighbors_ignored": int(case["n_ignored"])}
    if bool(case["query_given"]):
        kwargs["query"] = [int(q) for q in case["query"]]
    if bool(case["haystack_given"]):
        kwargs["haystack"] = [int(h) for h in case["haystack"]]
    return kwargs
