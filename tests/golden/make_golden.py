"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    cd /root/reference && PYTHONDONTWRITEBYTECODE=1 \
        PYTHONPATH=/root/repo/oracle:/root/repo python /root/repo/tests/golden/make_golden.py

``contact_map`` is imported from /root/reference (its version discovery needs that cwd) and
runs on the oracle's ``mdtraj`` stand-in (``oracle/mdtraj``), whose two numerical functions
are the C restatement ``oracle/cm_oracle.c``.  Every fixture stores the INPUTS (coordinates,
box, topology arrays, parameters) next to the reference's OUTPUTS, so that the tests can run
where /root/reference does not exist.
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

import mdtraj as md                      # noqa: E402  (the oracle stand-in)
import contact_map as ref                # noqa: E402  (the unmodified reference)
from contact_map_b200 import synth       # noqa: E402  (only the input generator)

warnings.simplefilter("ignore")
assert "oracle" in md.__file__, md.__file__
assert ref.__file__.startswith("/root/reference"), ref.__file__


def shim_topology(atom_residue, residue_chain, elements, residue_names):
    top = md.Topology()
    chains, residues = {}, {}
    for a, r in enumerate(atom_residue):
        c = int(residue_chain[r])
        if c not in chains:
            chains[c] = top.add_chain()
        if r not in residues:
            residues[r] = top.add_residue(residue_names[r], chains[c], resSeq=int(r) + 1)
        top.add_atom("X%d" % a, md.element.get_by_symbol(elements[a]), residues[r], serial=a + 1)
    return top


def shim_trajectory(xyz, box, top):
    """Stand-in trajectory.  Like MDTraj it stores float32 cell lengths/angles and derives the
    vectors from them, so the vectors the reference sees (also after its ``_atom_slice``
    copies lengths/angles, atom_indexer.py:9-14) are ``traj.unitcell_vectors``; that array,
    narrowed to float32 exactly as ``compute_neighborlist`` does, is what the fixture stores."""
    traj = md.Trajectory(xyz, top)
    if box is not None:
        traj.unitcell_vectors = box
    return traj


def counter_arrays(counter):
    keys = sorted(tuple(sorted(int(v) for v in k)) for k in counter.keys())
    pairs = np.array(keys, dtype=np.int64).reshape(-1, 2)
    vals = np.array([counter[frozenset(k)] for k in keys])
    return pairs, vals


def save_case(name, traj, top_arrays, query, haystack, cutoff, n_ignored, extra=None):
    atom_residue, residue_chain, elements, residue_names = top_arrays
    kwargs = {}
    if query is not None:
        kwargs["query"] = query
    if haystack is not None:
        kwargs["haystack"] = haystack
    freq = ref.ContactFrequency(traj, cutoff=cutoff, n_neighbors_ignored=n_ignored, **kwargs)
    ctraj = ref.ContactTrajectory(traj, cutoff=cutoff, n_neighbors_ignored=n_ignored, **kwargs)
    ap, ac = counter_arrays(freq._atom_contacts)
    rp, rc = counter_arrays(freq._residue_contacts)
    fa_ptr, fa_pairs, fr_ptr, fr_pairs = [0], [], [0], []
    for frame in ctraj:
        p, _ = counter_arrays(frame._atom_contacts)
        fa_pairs.append(p)
        fa_ptr.append(fa_ptr[-1] + len(p))
        p, _ = counter_arrays(frame._residue_contacts)
        fr_pairs.append(p)
        fr_ptr.append(fr_ptr[-1] + len(p))
    afreq_p, afreq_v = counter_arrays(freq.atom_contacts.counter)
    out = dict(
        xyz=np.asarray(traj.xyz, dtype=np.float32),
        atom_residue=np.asarray(atom_residue, dtype=np.int32),
        residue_chain=np.asarray(residue_chain, dtype=np.int32),
        atom_elements=np.array(elements), residue_names=np.array(residue_names),
        query=np.array(sorted(freq.query), dtype=np.int64),
        haystack=np.array(sorted(freq.haystack), dtype=np.int64),
        query_given=np.array(query is not None), haystack_given=np.array(haystack is not None),
        cutoff=np.array(cutoff, dtype=np.float64), n_ignored=np.array(n_ignored),
        atom_pairs=ap, atom_counts=ac, res_pairs=rp, res_counts=rc,
        atom_freq_pairs=afreq_p, atom_freq_values=np.asarray(afreq_v, dtype=np.float64),
        frame_atom_ptr=np.array(fa_ptr), frame_atom_pairs=np.concatenate(fa_pairs).reshape(-1, 2)
        if fa_pairs else np.empty((0, 2), np.int64),
        frame_res_ptr=np.array(fr_ptr), frame_res_pairs=np.concatenate(fr_pairs).reshape(-1, 2)
        if fr_pairs else np.empty((0, 2), np.int64),
        n_x=np.array(freq.atom_contacts.n_x.n), n_y=np.array(freq.atom_contacts.n_y.n),
        max_size=np.array(freq.atom_contacts.max_size),
        res_n_x=np.array(freq.residue_contacts.n_x.n), res_n_y=np.array(freq.residue_contacts.n_y.n),
        res_max_size=np.array(freq.residue_contacts.max_size),
        use_atom_slice=np.array(freq.use_atom_slice),
    )
    if traj._have_unitcell:
        out["box"] = np.asarray(traj.unitcell_vectors, dtype=np.float32)
    if extra:
        out.update(extra)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("%-28s atoms=%4d frames=%2d atom pairs=%5d residue pairs=%4d" % (
        name, traj.n_atoms, len(traj), len(ap), len(rp)))
    return freq


def pdb_cases():
    traj = md.load("/root/reference/contact_map/tests/trajectory.pdb")
    top = traj.topology
    atom_residue = [a.residue.index for a in top.atoms]
    residue_chain = [r.chain.index for r in top.residues]
    elements = [a.element.symbol for a in top.atoms]
    residue_names = [r.name for r in top.residues]
    arrays = (atom_residue, residue_chain, elements, residue_names)

    # min-dist goldens (test_min_dist.py)
    mdc = ref.MinimumDistanceCounter(traj, [4, 5], [6, 7])
    hist = np.array([[a.index, b.index] for a, b in mdc.atom_history])
    near = {}
    for frame, excluded, tag in ((0, None, "f0_res"), (4, None, "f4_res"), (4, {}, "f4_all")):
        na = ref.NearestAtoms(traj, cutoff=0.075, frame_number=frame, excluded=excluded)
        atoms = sorted(na.nearest)
        near["nearest_%s_atom" % tag] = np.array(atoms)
        near["nearest_%s_partner" % tag] = np.array([na.nearest[a] for a in atoms])
        near["nearest_%s_dist" % tag] = np.array([na.nearest_distance[a] for a in atoms], dtype=np.float32)
    extra = dict(mindist_query=np.array([4, 5]), mindist_haystack=np.array([6, 7]),
                 mindist_values=np.asarray(mdc.minimum_distances, dtype=np.float32),
                 mindist_pairs=np.asarray(mdc._min_pairs), mindist_history=hist, **near)

    f_all = save_case("pdb_c0075_n0", traj, arrays, None, None, 0.075, 0, extra)
    save_case("pdb_c0075_n1", traj, arrays, None, None, 0.075, 1)
    save_case("pdb_c045_n2", traj, arrays, None, None, 0.45, 2)
    save_case("pdb_c0075_n0_q01_h", traj, arrays, [0, 1], None, 0.075, 0)
    save_case("pdb_c0075_n0_q_sub", traj, arrays, [1, 4, 5], [4, 6, 7, 8], 0.075, 0)
    traj_nocell = md.load("/root/reference/contact_map/tests/trajectory.pdb")
    traj_nocell.unitcell_vectors = None
    save_case("pdb_c0075_n0_nocell", traj_nocell, arrays, None, None, 0.075, 0)

    # ContactDifference goldens (test_contact_map.py:586-715)
    f0 = ref.ContactFrequency(traj[0], cutoff=0.075, n_neighbors_ignored=0)
    f4 = ref.ContactFrequency(traj[4], cutoff=0.075, n_neighbors_ignored=0)
    out = {}
    for tag, diff in (("traj_minus_f0", f_all - f0), ("f4_minus_f0", f4 - f0), ("f0_minus_traj", f0 - f_all)):
        p, v = counter_arrays(diff.atom_contacts.counter)
        out[tag + "_atom_pairs"], out[tag + "_atom_values"] = p, np.asarray(v, dtype=np.float64)
        p, v = counter_arrays(diff.residue_contacts.counter)
        out[tag + "_res_pairs"], out[tag + "_res_values"] = p, np.asarray(v, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "pdb_differences.npz"), **out)

    # rolling-window goldens (test_contact_trajectory.py:288-348)
    ct = ref.ContactTrajectory(traj, cutoff=0.075, n_neighbors_ignored=0)
    out = {}
    for width, step in ((2, 1), (3, 2), (5, 1)):
        for w, cmap in enumerate(ref.RollingContactFrequency(ct, width=width, step=step)):
            p, v = counter_arrays(cmap._atom_contacts)
            out["w%d_s%d_%d_pairs" % (width, step, w)] = p
            out["w%d_s%d_%d_counts" % (width, step, w)] = v
            out["w%d_s%d_%d_nframes" % (width, step, w)] = np.array(cmap.n_frames)
        out["w%d_s%d_n" % (width, step)] = np.array(w + 1)
    np.savez_compressed(os.path.join(HERE, "pdb_rolling.npz"), **out)


def synthetic_cases():
    specs = [
        # name, n_atoms, box, cutoff, n_ignored, frames, q_frac, h_frac, chains
        ("syn_ortho_all", 120, "ortho", 0.45, 2, 4, None, None, 2),
        ("syn_ortho_sub", 150, "ortho", 0.40, 1, 4, 0.4, 0.6, 3),
        ("syn_tric_all", 140, "triclinic", 0.45, 2, 4, None, None, 2),
        ("syn_tric_sub", 160, "triclinic", 0.55, 0, 3, 0.5, 0.5, 3),
        ("syn_open_all", 100, "none", 0.45, 2, 3, None, None, 1),
        ("syn_open_sub", 130, "none", 0.60, 3, 3, 0.3, 1.0, 4),
        ("syn_ortho_n0", 90, "ortho", 0.35, 0, 3, None, None, 2),
    ]
    for k, (name, n, box_kind, cutoff, n_ign, frames, qf, hf, chains) in enumerate(specs):
        rng = np.random.default_rng(4200 + k)
        system = synth.random_system(rng, n, n_chains=chains, box_kind=box_kind, extent=1.3,
                                     h_fraction=0.25 if qf is None else 0.0)
        t = system.topology
        arrays = (t.atom_residue, t.residue_chain, t.atom_elements, t.residue_names)
        top = shim_topology(*arrays)
        xyz = system.frames(0, frames)
        vec = None if system.box is None else np.broadcast_to(system.box, (frames, 3, 3)).copy()
        traj = shim_trajectory(xyz, vec, top)
        query = None if qf is None else sorted(rng.choice(n, size=int(qf * n), replace=False).tolist())
        hay = None if hf is None else sorted(rng.choice(n, size=int(hf * n), replace=False).tolist())
        extra = {}
        if k in (0, 2, 4):
            q = list(range(0, 10))
            h = list(range(40, 90))
            mdc = ref.MinimumDistanceCounter(traj, q, h)
            extra.update(mindist_query=np.array(q), mindist_haystack=np.array(h),
                         mindist_values=np.asarray(mdc.minimum_distances, dtype=np.float32),
                         mindist_pairs=np.asarray(mdc._min_pairs))
            na = ref.NearestAtoms(traj, cutoff=cutoff, frame_number=1)
            atoms = sorted(na.nearest)
            extra.update(nearest_atom=np.array(atoms),
                         nearest_partner=np.array([na.nearest[a] for a in atoms]),
                         nearest_dist=np.array([na.nearest_distance[a] for a in atoms], dtype=np.float32))
        save_case(name, traj, arrays, query, hay, cutoff, n_ign, extra)


if __name__ == "__main__":
    pdb_cases()
    synthetic_cases()
