"""Short S1 run for ncu: a few launches of the frame kernel on device-resident frames.

    python tools/profile_s1.py [n_frames] [launches] [system]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from contact_map_b200 import synth                    # noqa: E402
from contact_map_b200.engine import get_engine        # noqa: E402


def main():
    n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    launches = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    which = sys.argv[3] if len(sys.argv) > 3 else "s1"
    system = {"s1": synth.kras_like, "s2": synth.binding_system,
              "s3": lambda: synth.solvated_triclinic(n_atoms=100000)}[which]()
    top = system.topology
    n = top.n_atoms
    engine = get_engine(0)
    atom_res, atom_chain = top.atom_residue_chain()
    role = np.zeros(n, dtype=np.uint8)
    role[np.arange(n) if system.query is None else system.query] |= 1
    role[np.arange(n) if system.haystack is None else system.haystack] |= 2
    if os.environ.get("CMB_LARGE_BATCH"):
        engine.set_option("large_batch", int(os.environ["CMB_LARGE_BATCH"]))
    engine.set_system(atom_res, atom_chain, role, 0.45, 2)
    dev = engine.device
    d_xyz = engine.synth_frames(torch.as_tensor(system.base, device=dev),
                                torch.as_tensor(system.groups.astype(np.int32), device=dev),
                                torch.as_tensor(system.box.reshape(9), device=dev), system.seed,
                                system.amplitude, True, 0, n_frames)
    d_box = torch.as_tensor(system.box, device=dev).reshape(1, 3, 3).expand(n_frames, 3, 3).contiguous()
    atom_t, res_t = engine.new_tables(hashed=True if os.environ.get("CMB_PROFILE_HASHED") else None)
    for _ in range(launches):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        engine.accumulate(d_xyz, d_box, atom_t, res_t)
        e1.record()
        torch.cuda.synchronize()
        print("%s: %d frames in %.3f ms -> %.0f frames/s" % (which, n_frames, e0.elapsed_time(e1),
                                                            n_frames / e0.elapsed_time(e1) * 1e3))
    print("nonzero atom pairs:", len(engine.export_nonzero(atom_t)[0]))


if __name__ == "__main__":
    main()
