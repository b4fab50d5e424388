"""CPU: the oracle against the golden vectors produced by the UNMODIFIED reference."""
import numpy as np
import pytest

from oracle import clib, refpath
from tests import goldens


@pytest.mark.parametrize("name", goldens.CASE_NAMES)
def test_refpath_matches_reference(name):
    case = goldens.load(name)
    atoms, residues = refpath.contact_frequency(
        case["xyz"], case.get("box"), case["atom_residue"], case["residue_chain"],
        case["query"], case["haystack"], float(case["cutoff"]), int(case["n_ignored"]))
    assert atoms == goldens.counter(case["atom_pairs"], case["atom_counts"])
    assert residues == goldens.counter(case["res_pairs"], case["res_counts"])


@pytest.mark.parametrize("name", goldens.CASE_NAMES)
@pytest.mark.parametrize("cells", [False, True])
def test_closed_form_matches_reference(name, cells):
    case = goldens.load(name)
    atoms, residues = refpath.closed_form_frequency(
        case["xyz"], case.get("box"), case["atom_residue"], case["residue_chain"],
        case["query"], case["haystack"], float(case["cutoff"]), int(case["n_ignored"]), cells=cells)
    assert atoms == goldens.counter(case["atom_pairs"], case["atom_counts"])
    assert residues == goldens.counter(case["res_pairs"], case["res_counts"])


@pytest.mark.parametrize("name", goldens.CASE_NAMES)
def test_per_frame_sets_match_reference(name):
    case = goldens.load(name)
    frames = refpath.contact_trajectory(
        case["xyz"], case.get("box"), case["atom_residue"], case["residue_chain"],
        case["query"], case["haystack"], float(case["cutoff"]), int(case["n_ignored"]))
    want_atoms = goldens.frame_sets(case["frame_atom_ptr"], case["frame_atom_pairs"])
    want_res = goldens.frame_sets(case["frame_res_ptr"], case["frame_res_pairs"])
    assert [f[0] for f in frames] == want_atoms
    assert [f[1] for f in frames] == want_res


def test_reference_known_answers_pdb():
    """Literal goldens of the reference's own tests (test_contact_map.py:19-39,207-214)."""
    case = goldens.load("pdb_c0075_n0")
    want = {(0, 8): 1, (0, 9): 1, (1, 4): 4, (1, 5): 1, (1, 8): 1, (1, 9): 1, (4, 6): 5, (4, 7): 2,
            (4, 8): 1, (5, 6): 5, (5, 7): 2, (5, 8): 1}
    got = {tuple(int(v) for v in p): int(c) for p, c in zip(case["atom_pairs"], case["atom_counts"])}
    assert got == want
    got = {tuple(int(v) for v in p): int(c) for p, c in zip(case["res_pairs"], case["res_counts"])}
    assert got == {(0, 2): 5, (0, 4): 1, (2, 3): 5, (2, 4): 1}
    # config 1 of BASELINE.json: cutoff 0.45, n_ignored 2 -> every pair > 2 residues apart, all frames
    case = goldens.load("pdb_c045_n2")
    got = {tuple(int(v) for v in p): int(c) for p, c in zip(case["atom_pairs"], case["atom_counts"])}
    want = {(a, b): 5 for a in (0, 1) for b in (6, 7, 8, 9)}
    want.update({(a, b): 5 for a in (2, 3) for b in (8, 9)})
    assert got == want
    # min-distance series of test_min_dist.py:129-135 and histories :115-116
    case = goldens.load("pdb_c0075_n0")
    assert case["mindist_values"] == pytest.approx([0.045276925690687, 0.045276925690687,
                                                    0.045276925690687, 0.05, 0.051], rel=1e-6)
    assert case["mindist_history"].tolist() == [[5, 6], [4, 6], [5, 6], [5, 7], [4, 6]]


@pytest.mark.parametrize("name", ["pdb_c0075_n0", "syn_ortho_all", "syn_tric_all", "syn_open_all"])
def test_min_dist_and_nearest(name):
    case = goldens.load(name)
    box = case.get("box")
    ortho = name != "syn_tric_all"
    vals, arg = clib.min_dist(case["xyz"], case["mindist_query"], case["mindist_haystack"], box, ortho)
    assert np.array_equal(vals, case["mindist_values"])
    assert np.array_equal(arg, case["mindist_pairs"])
    vals2, arg2 = refpath.minimum_distances(case["xyz"], box, case["mindist_query"].tolist(),
                                            case["mindist_haystack"].tolist(), ortho)
    assert np.array_equal(vals2, vals) and np.array_equal(arg2, arg)
    atom_res = case["atom_residue"]
    excluded = {i: set(np.nonzero(atom_res == atom_res[i])[0].tolist()) for i in range(len(atom_res))}
    if name.startswith("pdb"):
        frame, key = 4, "nearest_f4_res"
        cutoff = 0.075
        want = (case[key + "_atom"], case[key + "_partner"], case[key + "_dist"])
    else:
        frame, cutoff = 1, float(case["cutoff"])
        want = (case["nearest_atom"], case["nearest_partner"], case["nearest_dist"])
    near, dist = refpath.nearest_atoms(case["xyz"][frame], None if box is None else box[frame],
                                       cutoff, excluded, ortho)
    atoms = sorted(near)
    assert atoms == want[0].tolist()
    assert [near[a] for a in atoms] == want[1].tolist()
    assert np.array_equal(np.array([dist[a] for a in atoms], dtype=np.float32), want[2])


def _emulate_nl_d2(pi, pj, box):
    """numpy float32 re-derivation of the neighbour-list arithmetic, op by op."""
    f = np.float32
    d = (pj.astype(f) - pi.astype(f)).astype(f)
    if box is not None:
        b = box.astype(f)
        tri = any(b[r, c] != 0 for r in range(3) for c in range(3) if r != c)
        inv = (f(1.0) / np.array([b[0, 0], b[1, 1], b[2, 2]], dtype=f)).astype(f)
        if tri:
            for k in (2, 1, 0):
                s = np.floor(f(f(d[k] * inv[k]) + f(0.5))).astype(f)
                d = (d - (b[k] * s).astype(f)).astype(f)
        else:
            size = np.array([b[0, 0], b[1, 1], b[2, 2]], dtype=f)
            d = (d - (np.rint((d * inv).astype(f)).astype(f) * size).astype(f)).astype(f)
    sq = (d * d).astype(f)
    return f(f(sq[0] + sq[1]) + sq[2])


@pytest.mark.parametrize("kind", ["none", "ortho", "triclinic"])
def test_nl_arithmetic_matches_numpy_emulation(kind):
    rng = np.random.default_rng(7)
    box = None
    if kind == "ortho":
        box = np.diag([2.1, 2.6, 3.3]).astype(np.float32)
    elif kind == "triclinic":
        box = np.array([[2.5, 0, 0], [0.7, 2.2, 0], [-0.6, 0.5, 2.9]], dtype=np.float32)
    for _ in range(300):
        pi = rng.uniform(-6, 6, 3).astype(np.float32)
        pj = rng.uniform(-6, 6, 3).astype(np.float32)
        assert clib.nl_d2(pi, pj, box) == float(_emulate_nl_d2(pi, pj, box))


def test_cells_equal_brute_force_large():
    from contact_map_b200 import synth
    from tests.helpers import system_arrays
    system = synth.solvated_triclinic(n_atoms=6000, lattice=19)
    xyz = system.frames(0, 1)[0]
    res, chain, role = system_arrays(system)
    a = clib.contacts_frame(xyz, system.box, 0.45, res, chain, role, 2, cells=False)
    b = clib.contacts_frame(xyz, system.box, 0.45, res, chain, role, 2, cells=True)
    assert len(a) > 1000 and np.array_equal(a, b)


def test_neighborlist_cells_equal_brute_force():
    from contact_map_b200 import synth
    for system in (synth.kras_like(), synth.solvated_triclinic(n_atoms=3000, lattice=15)):
        xyz = system.frames(3, 1)[0]
        a = clib.neighborlist(xyz, system.box, 0.45, cells=False)
        b = clib.neighborlist(xyz, system.box, 0.45, cells=True)
        assert sum(len(v) for v in a) > 10000
        assert all(np.array_equal(u, v) for u, v in zip(a, b))


def test_boundary_report():
    # two atoms exactly at the cutoff along x: d2 == c2 in float32 -> contact, reported as boundary
    c = np.float32(0.45)
    xyz = np.array([[0, 0, 0], [c, 0, 0], [np.nextafter(c, np.float32(1)), 1.0, 0]], dtype=np.float32)
    xyz[2, 1] = 0.0
    res = np.array([0, 5, 9], dtype=np.int32)
    chain = np.zeros(3, dtype=np.int32)
    role = np.full(3, 3, dtype=np.uint8)
    pairs = clib.contacts_frame(xyz, None, 0.45, res, chain, role, 2)
    assert pairs.tolist() == [[0, 1], [1, 2]]
    a, b, d2 = clib.boundary_pairs(xyz, None, 0.45, res, chain, role, 2, ulps=1)
    assert list(zip(a.tolist(), b.tolist())) == [(0, 1)]
    # (c + ulp(c))^2 is 2 float steps above c^2: not a contact, reported from ulps=2 on
    a, b, d2 = clib.boundary_pairs(xyz, None, 0.45, res, chain, role, 2, ulps=2)
    assert list(zip(a.tolist(), b.tolist())) == [(0, 1), (0, 2)]
