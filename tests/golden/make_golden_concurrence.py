"""Golden vectors for the contact concurrences, from the UNMODIFIED reference.

    cd /root/reference && PYTHONDONTWRITEBYTECODE=1 \
        PYTHONPATH=/root/repo/oracle:/root/repo python /root/repo/tests/golden/make_golden_concurrence.py

Same set-up as make_golden.py (reference Python over the oracle's mdtraj stand-in).  Stores the
inputs (coordinates, box, topology arrays) next to the reference's labels and boolean tables.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.path.insert(0, HERE)

import mdtraj as md                                  # noqa: E402  (the oracle stand-in)
import contact_map as ref                            # noqa: E402  (the unmodified reference)
from contact_map.concurrence import AtomContactConcurrence, ResidueContactConcurrence   # noqa: E402
from contact_map_b200 import synth                   # noqa: E402
from make_golden import shim_topology, shim_trajectory                                   # noqa: E402

warnings.simplefilter("ignore")


def top_arrays(top):
    return dict(atom_residue=np.array([a.residue.index for a in top.atoms], dtype=np.int32),
                residue_chain=np.array([r.chain.index for r in top.residues], dtype=np.int32),
                atom_elements=np.array([a.element.symbol for a in top.atoms]),
                atom_names=np.array([a.name for a in top.atoms]),
                residue_names=np.array([r.name for r in top.residues]),
                residue_resseq=np.array([r.resSeq for r in top.residues], dtype=np.int64))


def save(name, traj, freq, cutoff):
    out = top_arrays(traj.topology)
    out["xyz"] = np.asarray(traj.xyz, dtype=np.float32)
    if traj._have_unitcell:
        out["box"] = np.asarray(traj.unitcell_vectors, dtype=np.float32)
    out["cutoff"] = np.array(cutoff, dtype=np.float64)
    out["query"] = np.array(sorted(freq.query), dtype=np.int64)
    out["haystack"] = np.array(sorted(freq.haystack), dtype=np.int64)
    out["freq_cutoff"] = np.array(freq.cutoff)
    out["freq_n_ignored"] = np.array(freq.n_neighbors_ignored)
    atom_mc = freq.atom_contacts.most_common()
    res_mc = freq.residue_contacts.most_common()
    out["atom_contact_pairs"] = np.array([[c[0][0].index, c[0][1].index] for c in atom_mc],
                                         dtype=np.int64).reshape(-1, 2)
    out["res_contact_pairs"] = np.array([[c[0][0].index, c[0][1].index] for c in res_mc],
                                        dtype=np.int64).reshape(-1, 2)
    atom_c = AtomContactConcurrence(traj, atom_mc, cutoff=cutoff)
    heavy = ResidueContactConcurrence(traj, res_mc, cutoff=cutoff)
    every = ResidueContactConcurrence(traj, res_mc, cutoff=cutoff, select="")
    out["atom_labels"] = np.array(atom_c.labels)
    out["atom_values"] = np.array(atom_c.values, dtype=bool).reshape(len(atom_c.labels), -1)
    out["res_labels"] = np.array(heavy.labels)
    out["res_values_heavy"] = np.array(heavy.values, dtype=bool).reshape(len(heavy.labels), -1)
    out["res_values_all"] = np.array(every.values, dtype=bool).reshape(len(every.labels), -1)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("%-26s frames=%d atom contacts=%d (true %d) residue contacts=%d (heavy true %d, all true %d)" % (
        name, len(traj), len(atom_c.labels), out["atom_values"].sum(), len(heavy.labels),
        out["res_values_heavy"].sum(), out["res_values_all"].sum()))


def main():
    # the reference's own fixture and parameters (test_concurrence.py:15-22)
    traj = md.load("/root/reference/contact_map/tests/concurrence.pdb")
    query = traj.topology.select("resSeq 3")
    haystack = traj.topology.select("resSeq 1 to 2")
    freq = ref.ContactFrequency(traj, query, haystack, cutoff=0.051, n_neighbors_ignored=0)
    save("conc_pdb", traj, freq, 0.051)
    # seeded synthetic systems: periodic cells, H atoms, several chains
    for k, (name, n, box_kind, cutoff) in enumerate([("conc_syn_ortho", 90, "ortho", 0.45),
                                                     ("conc_syn_tric", 110, "triclinic", 0.40),
                                                     ("conc_syn_open", 80, "none", 0.50)]):
        rng = np.random.default_rng(5200 + k)
        system = synth.random_system(rng, n, n_chains=2, box_kind=box_kind, extent=1.3, h_fraction=0.3)
        t = system.topology
        top = shim_topology(t.atom_residue, t.residue_chain, t.atom_elements, t.residue_names)
        frames = 6
        xyz = system.frames(0, frames)
        vec = None if system.box is None else np.broadcast_to(system.box, (frames, 3, 3)).copy()
        traj = shim_trajectory(xyz, vec, top)
        freq = ref.ContactFrequency(traj, cutoff=cutoff, n_neighbors_ignored=2)
        save(name, traj, freq, cutoff)


if __name__ == "__main__":
    main()
