"""Bit-reproducible synthetic trajectories.

Coordinates are generated without transcendental functions so host (numpy) and
device (``cmb_synth_frames``) produce identical bits:

    u(seed, frame, atom, axis) = (splitmix64(seed ^ frame*P1 ^ atom*P2 ^ axis*P3) >> 40) * 2^-24
    x = base[atom][axis] + A * (2u - 1)              (fp32, every op rounded separately)

and, "molecules not made whole", every residue gets a per-frame integer image
shift m in {-1, 0, +1} per box vector (hash with axis+3), added as m*box_vector
(fp32 mul then add, a, b, c in turn), so raw deltas span +-2 box lengths while
minimum-image distances are unchanged up to rounding.
"""
import numpy as np

from .trajectory import Topology, Trajectory

P1 = np.uint64(0xD6E8FEB86659FD93)
P2 = np.uint64(0xA24BAED4963EE407)
P3 = np.uint64(0x9FB21C651E98DF25)


def splitmix64(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def make_frames(base, groups, box, seed, amplitude, image_shifts, frame0, n_frames):
    """Host generator.  base float32 [N,3]; groups int [N] (residue of each atom);
    box float32 [3,3] or None.  Returns float32 [n_frames, N, 3]."""
    base = np.ascontiguousarray(base, dtype=np.float32)
    n = base.shape[0]
    amp = np.float32(amplitude)
    frames = np.arange(frame0, frame0 + n_frames, dtype=np.uint64)
    atoms = np.arange(n, dtype=np.uint64)
    out = np.empty((n_frames, n, 3), dtype=np.float32)
    with np.errstate(over="ignore"):
        fterm = (frames * P1)[:, None]
        aterm = (atoms * P2)[None, :]
        for ax in range(3):
            h = splitmix64(np.uint64(seed) ^ fterm ^ aterm ^ (np.uint64(ax) * P3))
            u = (h >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)
            w = (np.float32(2.0) * u) - np.float32(1.0)
            out[:, :, ax] = base[None, :, ax] + amp * w
        if image_shifts and box is not None:
            bx = np.ascontiguousarray(box, dtype=np.float32).reshape(3, 3)
            gterm = (np.asarray(groups, dtype=np.uint64) * P2)[None, :]
            for v in range(3):
                h = splitmix64(np.uint64(seed) ^ fterm ^ gterm ^ (np.uint64(v + 3) * P3))
                m = (((h >> np.uint64(40)) * np.uint64(3)) >> np.uint64(24)).astype(np.int64) - 1
                fm = m.astype(np.float32)
                for ax in range(3):
                    out[:, :, ax] = out[:, :, ax] + fm * bx[v, ax]
    return out


# ---------------------------------------------------------------------------------
# named systems
# ---------------------------------------------------------------------------------
class SyntheticSystem(object):
    """Topology + generator parameters of one named workload."""

    def __init__(self, name, topology, base, box, seed, amplitude, image_shifts=True,
                 query=None, haystack=None):
        self.name = name
        self.topology = topology
        self.base = np.ascontiguousarray(base, dtype=np.float32)
        self.box = None if box is None else np.ascontiguousarray(box, dtype=np.float32)
        self.seed = int(seed)
        self.amplitude = float(amplitude)
        self.image_shifts = bool(image_shifts)
        self.query = query
        self.haystack = haystack

    @property
    def n_atoms(self):
        return self.base.shape[0]

    @property
    def groups(self):
        return self.topology.atom_residue

    def frames(self, frame0, n_frames):
        return make_frames(self.base, self.groups, self.box, self.seed, self.amplitude,
                           self.image_shifts, frame0, n_frames)

    def trajectory(self, n_frames, frame0=0):
        cell = None if self.box is None else np.broadcast_to(self.box, (n_frames, 3, 3)).copy()
        return Trajectory(self.frames(frame0, n_frames), self.topology, cell)


def _globule_centres(n_res, spacing):
    """Residue centres on a boustrophedon walk through a cubic lattice (compact globule)."""
    side = int(np.ceil(n_res ** (1.0 / 3.0)))
    pts = []
    for z in range(side):
        ys = range(side) if z % 2 == 0 else range(side - 1, -1, -1)
        for yi, y in enumerate(ys):
            xs = range(side) if (yi + z * side) % 2 == 0 else range(side - 1, -1, -1)
            for x in xs:
                pts.append((x, y, z))
    return np.array(pts[:n_res], dtype=np.float64) * spacing


def kras_like(n_res=270, atoms_per_res=10, seed=1001, box=(7.0, 7.5, 8.0), amplitude=0.03):
    """S1: KRAS-sized protein, 270 residues x 10 heavy atoms, one chain, orthorhombic box."""
    rng = np.random.default_rng(seed)
    centres = _globule_centres(n_res, 0.57)
    centres += (np.array(box) - centres.max(axis=0)) / 2.0 - centres.min(axis=0) / 2.0
    offsets = rng.normal(0.0, 0.12, size=(n_res, atoms_per_res, 3))
    base = (centres[:, None, :] + offsets).reshape(-1, 3).astype(np.float32)
    n_atoms = n_res * atoms_per_res
    top = Topology(np.repeat(np.arange(n_res), atoms_per_res), np.zeros(n_res, dtype=np.int32),
                   atom_names=["C%d" % (i % atoms_per_res) for i in range(n_atoms)],
                   atom_elements=["C"] * n_atoms, residue_names=["ALA"] * n_res)
    cell = np.diag(np.array(box, dtype=np.float32))
    return SyntheticSystem("S1-kras-like", top, base, cell, seed, amplitude)


def binding_system(partner_res=60, partner_atoms_per_res=10, seed=1002, ligand=False):
    """S2: receptor = S1 protein (chain 0) + partner chain / ligand (chain 1) at its surface."""
    rec = kras_like(seed=seed)
    rng = np.random.default_rng(seed + 7)
    if ligand:
        partner_res, partner_atoms_per_res = 1, 30
    n_p = partner_res * partner_atoms_per_res
    lo, hi = rec.base.min(axis=0), rec.base.max(axis=0)
    anchor = np.array([hi[0] + 0.25, (lo[1] + hi[1]) / 2, (lo[2] + hi[2]) / 2])
    if ligand:
        pbase = anchor + rng.normal(0.0, 0.25, size=(n_p, 3))
    else:
        centres = _globule_centres(partner_res, 0.57)
        centres = centres - centres.mean(axis=0) + anchor + np.array([0.9, 0.0, 0.0])
        pbase = (centres[:, None, :] + rng.normal(0.0, 0.12, size=(partner_res, partner_atoms_per_res, 3))
                 ).reshape(-1, 3)
    base = np.concatenate([rec.base, pbase.astype(np.float32)], axis=0)
    n_rec_res = rec.topology.n_residues
    atom_res = np.concatenate([rec.topology.atom_residue,
                               n_rec_res + np.repeat(np.arange(partner_res), partner_atoms_per_res)])
    res_chain = np.concatenate([np.zeros(n_rec_res, dtype=np.int32), np.ones(partner_res, dtype=np.int32)])
    top = Topology(atom_res, res_chain, atom_elements=["C"] * base.shape[0],
                   residue_names=["ALA"] * n_rec_res + (["LIG"] if ligand else ["GLY"] * partner_res))
    n_rec = rec.n_atoms
    # larger jitter so the interface contacts fluctuate from frame to frame
    return SyntheticSystem("S2-ligand" if ligand else "S2-partner", top, base, rec.box, seed, 0.06,
                           query=np.arange(n_rec, n_rec + n_p), haystack=np.arange(n_rec))


def solvated_triclinic(n_atoms=100000, seed=1003, amplitude=0.03, lattice=47):
    """S3: atoms on the first n_atoms sites of a lattice^3 fractional lattice of a reduced
    triclinic cell; first 2700 sites form a 270x10 'protein' (chain 0), the rest 3-atom
    'solvent' residues (chain 1)."""
    rng = np.random.default_rng(seed)
    # the cell shrinks with the lattice so the site density (about 116 atoms/nm^3) is fixed
    cell = np.array([[10.2, 0.0, 0.0], [3.4, 9.6, 0.0], [-3.4, 3.2, 9.1]], dtype=np.float64)
    cell = cell * (lattice / 47.0)
    idx = np.arange(n_atoms)
    frac = np.stack([(idx % lattice), (idx // lattice) % lattice, idx // (lattice * lattice)],
                    axis=1).astype(np.float64)
    frac = (frac + 0.5) / lattice
    base = frac @ cell + rng.uniform(-0.05, 0.05, size=(n_atoms, 3))
    n_prot = min(2700, n_atoms)
    prot_res = np.repeat(np.arange(n_prot // 10 + (1 if n_prot % 10 else 0)), 10)[:n_prot]
    n_prot_res = int(prot_res.max()) + 1 if n_prot else 0
    n_solv = n_atoms - n_prot
    solv_res = n_prot_res + np.arange(n_solv) // 3
    atom_res = np.concatenate([prot_res, solv_res]).astype(np.int32)
    n_res = int(atom_res.max()) + 1
    res_chain = np.zeros(n_res, dtype=np.int32)
    res_chain[n_prot_res:] = 1
    top = Topology(atom_res, res_chain, atom_elements=["C"] * n_atoms,
                   residue_names=["ALA"] * n_prot_res + ["SOL"] * (n_res - n_prot_res))
    return SyntheticSystem("S3-solvated-triclinic", top, base.astype(np.float32),
                           cell.astype(np.float32), seed, amplitude,
                           query=np.arange(n_atoms), haystack=np.arange(n_atoms))


def random_system(rng, n_atoms, n_chains=2, box_kind="ortho", extent=2.0, h_fraction=0.2):
    """Small random system for parity tests: random residue sizes, chains, H atoms."""
    sizes = []
    while sum(sizes) < n_atoms:
        sizes.append(int(rng.integers(1, 7)))
    sizes[-1] -= sum(sizes) - n_atoms
    sizes = [s for s in sizes if s > 0]
    n_res = len(sizes)
    atom_res = np.repeat(np.arange(n_res), sizes)
    cuts = np.sort(rng.choice(np.arange(1, n_res), size=min(n_chains - 1, max(n_res - 1, 0)),
                              replace=False)) if n_res > 1 and n_chains > 1 else np.array([], dtype=int)
    res_chain = np.searchsorted(cuts, np.arange(n_res), side="right").astype(np.int32)
    elements = ["H" if rng.random() < h_fraction else "C" for _ in range(n_atoms)]
    top = Topology(atom_res, res_chain, atom_elements=elements, residue_names=["ALA"] * n_res)
    if box_kind == "ortho":
        box = np.diag(rng.uniform(extent, extent * 1.3, size=3)).astype(np.float32)
    elif box_kind == "triclinic":
        a = rng.uniform(extent, extent * 1.3, size=3)
        box = np.array([[a[0], 0, 0], [0.3 * a[0], a[1], 0], [-0.25 * a[0], 0.2 * a[1], a[2]]],
                       dtype=np.float32)
    else:
        box = None
    base = rng.uniform(0.0, extent, size=(n_atoms, 3)).astype(np.float32)
    return SyntheticSystem("random", top, base, box, int(rng.integers(1, 2 ** 31)), 0.05,
                           image_shifts=box is not None)
