"""Minimal trajectory / topology containers.

The reference takes ``mdtraj.Trajectory`` objects (third-party, not available
here).  The contact classes of this package are duck-typed: they accept any
object offering the small part of that interface the reference touches
(``len()``, ``.topology`` with ``atom(i)``, ``residue(i)``,
``atoms``, ``residues``, ``n_atoms``, ``n_residues``, ``select``; ``.xyz`` float32
``[F, N, 3]`` in nm; ``.unitcell_vectors`` ``[F, 3, 3]`` or None; int/slice
indexing).  This module provides a self-contained, array-backed implementation of
exactly that interface so the package works without MDTraj (synthetic workloads,
tests, numpy/PDB input).
"""
import numpy as np

_WATER = {"H2O", "HHO", "OHH", "HOH", "OH2", "SOL", "WAT", "TIP", "TIP2", "TIP3", "TIP4"}


class Atom(object):
    __slots__ = ("_top", "index")

    def __init__(self, top, index):
        self._top = top
        self.index = int(index)

    @property
    def name(self):
        return self._top.atom_names[self.index]

    @property
    def element_symbol(self):
        return self._top.atom_elements[self.index]

    @property
    def serial(self):
        return int(self._top.atom_serials[self.index])

    @property
    def residue(self):
        return Residue(self._top, self._top.atom_residue[self.index])

    def __eq__(self, other):
        return isinstance(other, Atom) and other.index == self.index and other._top == self._top

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash(("atom", self.index))

    def __str__(self):
        return "%s-%s" % (self.residue, self.name)

    __repr__ = __str__


class Residue(object):
    __slots__ = ("_top", "index")

    def __init__(self, top, index):
        self._top = top
        self.index = int(index)

    @property
    def name(self):
        return self._top.residue_names[self.index]

    @property
    def resSeq(self):
        return int(self._top.residue_resseq[self.index])

    @property
    def chain(self):
        return Chain(self._top, self._top.residue_chain[self.index])

    @property
    def atoms(self):
        lo, hi = self._top._res_atom_range(self.index)
        return (Atom(self._top, i) for i in range(lo, hi))

    @property
    def n_atoms(self):
        lo, hi = self._top._res_atom_range(self.index)
        return hi - lo

    @property
    def is_water(self):
        return self.name in _WATER

    def __eq__(self, other):
        return isinstance(other, Residue) and other.index == self.index and other._top == self._top

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash(("residue", self.index))

    def __str__(self):
        return "%s%d" % (self.name, self.resSeq)

    __repr__ = __str__


class Chain(object):
    __slots__ = ("_top", "index")

    def __init__(self, top, index):
        self._top = top
        self.index = int(index)

    @property
    def residues(self):
        idx = np.nonzero(self._top.residue_chain == self.index)[0]
        return (Residue(self._top, i) for i in idx)

    def __eq__(self, other):
        return isinstance(other, Chain) and other.index == self.index and other._top == self._top

    def __hash__(self):
        return hash(("chain", self.index))


class Topology(object):
    """Array-backed topology: atoms are contiguous per residue, residues per chain."""

    def __init__(self, atom_residue, residue_chain, atom_names=None, atom_elements=None,
                 residue_names=None, residue_resseq=None, atom_serials=None):
        self.atom_residue = np.ascontiguousarray(atom_residue, dtype=np.int32)
        self.residue_chain = np.ascontiguousarray(residue_chain, dtype=np.int32)
        n_atoms, n_res = self.atom_residue.shape[0], self.residue_chain.shape[0]
        if n_atoms and (np.any(np.diff(self.atom_residue) < 0) or self.atom_residue.max() >= n_res):
            raise ValueError("atoms must be ordered by residue index")
        self.atom_names = list(atom_names) if atom_names is not None else ["X"] * n_atoms
        self.atom_elements = list(atom_elements) if atom_elements is not None else ["C"] * n_atoms
        self.residue_names = list(residue_names) if residue_names is not None else ["RES"] * n_res
        self.residue_resseq = (np.asarray(residue_resseq, dtype=np.int64) if residue_resseq is not None
                               else np.arange(1, n_res + 1, dtype=np.int64))
        self.atom_serials = (np.asarray(atom_serials, dtype=np.int64) if atom_serials is not None
                             else np.arange(1, n_atoms + 1, dtype=np.int64))
        self._res_first = np.searchsorted(self.atom_residue, np.arange(n_res + 1), side="left")

    # -- sizes ----------------------------------------------------------
    @property
    def n_atoms(self):
        return int(self.atom_residue.shape[0])

    @property
    def n_residues(self):
        return int(self.residue_chain.shape[0])

    @property
    def n_chains(self):
        return int(self.residue_chain.max()) + 1 if self.n_residues else 0

    def _res_atom_range(self, res_index):
        return int(self._res_first[res_index]), int(self._res_first[res_index + 1])

    # -- object views ---------------------------------------------------
    def atom(self, index):
        if not 0 <= index < self.n_atoms:
            raise IndexError(index)
        return Atom(self, index)

    def residue(self, index):
        if not 0 <= index < self.n_residues:
            raise IndexError(index)
        return Residue(self, index)

    def chain(self, index):
        return Chain(self, index)

    @property
    def atoms(self):
        return (Atom(self, i) for i in range(self.n_atoms))

    @property
    def residues(self):
        return (Residue(self, i) for i in range(self.n_residues))

    @property
    def chains(self):
        return (Chain(self, i) for i in range(self.n_chains))

    # -- the arrays the engine wants -------------------------------------
    def atom_residue_chain(self):
        """(atom -> residue index, atom -> chain index) as int32 arrays."""
        return self.atom_residue, self.residue_chain[self.atom_residue]

    # -- selection (the subset of the MDTraj DSL the reference uses) --------
    def select(self, expression):
        expr = " ".join(expression.split())
        elements = np.array(self.atom_elements)
        is_water = np.array([name in _WATER for name in self.residue_names])[self.atom_residue]
        if expr in ("all", "everything"):
            mask = np.ones(self.n_atoms, dtype=bool)
        elif expr == "not water and symbol != 'H'":
            mask = (~is_water) & (elements != "H")
        elif expr == "not water":
            mask = ~is_water
        elif expr == "water":
            mask = is_water
        elif expr.startswith("resSeq "):
            parts = expr.split()
            if len(parts) == 2:
                mask = (self.residue_resseq == int(parts[1]))[self.atom_residue]
            elif len(parts) == 4 and parts[2] == "to":
                rs = self.residue_resseq
                mask = ((rs >= int(parts[1])) & (rs <= int(parts[3])))[self.atom_residue]
            else:
                raise ValueError("unsupported selection %r" % expression)
        elif expr.startswith("resid "):
            mask = np.isin(self.atom_residue, [int(p) for p in expr.split()[1:]])
        elif expr.startswith("chainid "):
            mask = np.isin(self.residue_chain[self.atom_residue], [int(p) for p in expr.split()[1:]])
        elif expr.startswith("symbol == "):
            mask = elements == expr.split("==")[1].strip().strip("'\"")
        elif expr.startswith("symbol != "):
            mask = elements != expr.split("!=")[1].strip().strip("'\"")
        else:
            raise ValueError("unsupported selection %r (this container implements only the "
                             "expressions the contact-map path uses)" % expression)
        return np.nonzero(mask)[0]

    # -- copies / comparison ---------------------------------------------
    def copy(self):
        return Topology(self.atom_residue.copy(), self.residue_chain.copy(), list(self.atom_names),
                        list(self.atom_elements), list(self.residue_names),
                        self.residue_resseq.copy(), self.atom_serials.copy())

    def subset(self, atom_indices):
        idx = np.asarray(sorted(int(i) for i in atom_indices), dtype=np.int64)
        old_res = self.atom_residue[idx]
        kept_res, new_res = np.unique(old_res, return_inverse=True)
        kept_chain, new_chain = np.unique(self.residue_chain[kept_res], return_inverse=True)
        return Topology(new_res, new_chain, [self.atom_names[i] for i in idx],
                        [self.atom_elements[i] for i in idx],
                        [self.residue_names[r] for r in kept_res], self.residue_resseq[kept_res],
                        self.atom_serials[idx])

    # -- data frame round trip (the columns of mdtraj.Topology.to_dataframe) ---------------
    def to_dataframe(self):
        """(atoms DataFrame, bonds array): what the reference serialises in ``to_dict``
        (contact_map.py:278-283).  This container holds no bonds: the array is empty."""
        import pandas as pd
        res = self.atom_residue
        atoms = pd.DataFrame({
            "serial": self.atom_serials,
            "name": list(self.atom_names),
            "element": list(self.atom_elements),
            "resSeq": self.residue_resseq[res],
            "resName": [self.residue_names[r] for r in res],
            "chainID": self.residue_chain[res],
            "segmentID": [""] * self.n_atoms,
        })
        return atoms, np.zeros((0, 4))

    @classmethod
    def from_dataframe(cls, atoms, bonds=None):
        """Inverse of :meth:`to_dataframe`; a new residue starts where (chainID, resSeq, resName,
        segmentID) changes, as in MDTraj.  Bonds are not kept."""
        n = len(atoms)
        chain = np.asarray(atoms["chainID"])
        resseq = np.asarray(atoms["resSeq"])
        resname = [str(v) for v in atoms["resName"]]
        seg = [("" if v is None else str(v)) for v in atoms["segmentID"]] if "segmentID" in atoms.columns \
            else [""] * n
        atom_res = np.zeros(n, dtype=np.int32)
        res_chain, res_names, res_seq = [], [], []
        chain_ids = {}
        prev = None
        for i in range(n):
            key = (chain[i], resseq[i], resname[i], seg[i])
            if key != prev:
                res_chain.append(chain_ids.setdefault(chain[i], len(chain_ids)))
                res_names.append(resname[i])
                res_seq.append(int(resseq[i]))
                prev = key
            atom_res[i] = len(res_chain) - 1
        serial = np.asarray(atoms["serial"])
        serials = None if serial.dtype == object or np.any(np.isnan(serial.astype(np.float64))) \
            else serial.astype(np.int64)
        return cls(atom_res, np.asarray(res_chain, dtype=np.int32), [str(v) for v in atoms["name"]],
                   [("" if v is None else str(v)) for v in atoms["element"]], res_names,
                   np.asarray(res_seq, dtype=np.int64), serials)

    def _signature(self):
        return (self.atom_residue.tobytes(), self.residue_chain.tobytes(), tuple(self.atom_names),
                tuple(self.atom_elements), tuple(self.residue_names), self.residue_resseq.tobytes())

    def __eq__(self, other):
        if self is other:
            return True
        if not isinstance(other, Topology):
            return False
        return self._signature() == other._signature()

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash(self._signature())

    def __repr__(self):
        return "<Topology: %d chains, %d residues, %d atoms>" % (self.n_chains, self.n_residues,
                                                                 self.n_atoms)


class Trajectory(object):
    """xyz float32 [F, N, 3] (nm) + optional unit-cell row vectors [F, 3, 3] + topology."""

    def __init__(self, xyz, topology, unitcell_vectors=None):
        xyz = np.asarray(xyz, dtype=np.float32)
        if xyz.ndim == 2:
            xyz = xyz[np.newaxis]
        if xyz.ndim != 3 or xyz.shape[2] != 3:
            raise ValueError("xyz must have shape [n_frames, n_atoms, 3]")
        if topology is not None and topology.n_atoms != xyz.shape[1]:
            raise ValueError("topology has %d atoms but xyz has %d" % (topology.n_atoms, xyz.shape[1]))
        self.xyz = np.ascontiguousarray(xyz)
        self.topology = topology
        self.unitcell_vectors = unitcell_vectors

    @property
    def unitcell_vectors(self):
        return self._unitcell_vectors

    @unitcell_vectors.setter
    def unitcell_vectors(self, vectors):
        if vectors is None:
            self._unitcell_vectors = None
            return
        vectors = np.asarray(vectors, dtype=np.float32)
        if vectors.ndim == 2:
            vectors = np.broadcast_to(vectors, (self.xyz.shape[0], 3, 3))
        if vectors.shape != (self.xyz.shape[0], 3, 3):
            raise ValueError("unitcell_vectors must have shape [n_frames, 3, 3]")
        self._unitcell_vectors = np.ascontiguousarray(vectors)

    @property
    def _have_unitcell(self):
        return self._unitcell_vectors is not None

    @property
    def top(self):
        return self.topology

    @property
    def n_frames(self):
        return self.xyz.shape[0]

    @property
    def n_atoms(self):
        return self.xyz.shape[1]

    def __len__(self):
        return self.n_frames

    def __getitem__(self, key):
        if isinstance(key, (int, np.integer)):
            key = slice(int(key), int(key) + 1) if key != -1 else slice(-1, None)
        cell = None if self._unitcell_vectors is None else self._unitcell_vectors[key]
        return Trajectory(self.xyz[key], self.topology, cell)

    def atom_slice(self, atom_indices):
        idx = np.asarray(sorted(int(i) for i in atom_indices), dtype=np.int64)
        return Trajectory(np.ascontiguousarray(self.xyz[:, idx]), self.topology.subset(idx),
                          self._unitcell_vectors)

    def join(self, other):
        others = other if isinstance(other, (list, tuple)) else [other]
        return join([self] + list(others))

    def __repr__(self):
        return "<Trajectory: %d frames, %d atoms%s>" % (
            self.n_frames, self.n_atoms, "" if self._unitcell_vectors is None else ", unit cell")


def join(trajs):
    trajs = list(trajs)
    xyz = np.concatenate([t.xyz for t in trajs], axis=0)
    cell = None
    if all(t.unitcell_vectors is not None for t in trajs):
        cell = np.concatenate([t.unitcell_vectors for t in trajs], axis=0)
    return Trajectory(xyz, trajs[0].topology, cell)


def _cell_vectors(a, b, c, alpha, beta, gamma):
    """Reduced-form box vectors from lengths (nm) / angles (deg): MDTraj's published recipe."""
    alpha, beta, gamma = (np.deg2rad(v) for v in (alpha, beta, gamma))
    av = np.array([a, 0.0, 0.0])
    bv = np.array([b * np.cos(gamma), b * np.sin(gamma), 0.0])
    cx = c * np.cos(beta)
    cy = c * (np.cos(alpha) - np.cos(beta) * np.cos(gamma)) / np.sin(gamma)
    cz = np.sqrt(c * c - cx * cx - cy * cy)
    cv = np.array([cx, cy, cz])
    out = np.array([av, bv, cv])
    out[np.abs(out) < 1e-6] = 0.0
    return out.astype(np.float32)


def load_pdb(filename):
    """Read a (multi-MODEL) PDB file: Angstrom -> nm as float32(float64 * 0.1)."""
    frames, cells, current = [], [], []
    cell = None
    atom_residue, residue_chain = [], []
    atom_names, atom_elements, residue_names, residue_resseq, atom_serials = [], [], [], [], []
    chain_ids = {}
    last_key = None
    first_done = False

    def finish():
        nonlocal current, first_done
        if current:
            frames.append(current)
            cells.append(cell)
            first_done = True
        current = []

    with open(filename) as fh:
        for line in fh:
            rec = line[:6]
            if rec == "CRYST1":
                cell = [float(line[6:15]), float(line[15:24]), float(line[24:33]),
                        float(line[33:40]), float(line[40:47]), float(line[47:54])]
            elif rec.startswith("MODEL"):
                finish()
            elif rec == "ENDMDL":
                finish()
            elif rec in ("ATOM  ", "HETATM"):
                current.append((float(line[30:38]), float(line[38:46]), float(line[46:54])))
                if not first_done:
                    chain_id = line[21]
                    if chain_id not in chain_ids:
                        chain_ids[chain_id] = len(chain_ids)
                        last_key = None
                    key = (chain_id, int(line[22:26]), line[17:20].strip())
                    if key != last_key:
                        residue_chain.append(chain_ids[chain_id])
                        residue_names.append(key[2])
                        residue_resseq.append(key[1])
                        last_key = key
                    atom_residue.append(len(residue_chain) - 1)
                    name = line[12:16].strip()
                    atom_names.append(name)
                    symbol = line[76:78].strip() if len(line) > 76 else ""
                    atom_elements.append(symbol if symbol else name[:1])
                    atom_serials.append(int(line[6:11]))
            elif rec.startswith("TER"):
                last_key = None
    finish()
    xyz = (np.array(frames, dtype=np.float64) * 0.1).astype(np.float32)
    topology = Topology(atom_residue, residue_chain, atom_names, atom_elements, residue_names,
                        residue_resseq, atom_serials)
    vectors = None
    if cells and all(c is not None for c in cells):
        vectors = np.stack([_cell_vectors(c[0] * 0.1, c[1] * 0.1, c[2] * 0.1, c[3], c[4], c[5])
                            for c in cells])
    return Trajectory(xyz, topology, vectors)


def _topology_from_npz(data):
    return Topology(data["atom_residue"], data["residue_chain"],
                    atom_names=[str(s) for s in data["atom_names"]] if "atom_names" in data else None,
                    atom_elements=[str(s) for s in data["atom_elements"]] if "atom_elements" in data else None,
                    residue_names=[str(s) for s in data["residue_names"]] if "residue_names" in data else None,
                    residue_resseq=data["residue_resseq"] if "residue_resseq" in data else None)


def save_npy(filename, trajectory):
    """Write ``trajectory`` in the chunk-readable layout :func:`load` memory-maps:

    * ``<filename>`` (``.npy``): the coordinates, float32 ``[n_frames, n_atoms, 3]`` in nm, plain
      NumPy format -- any writer of ``.npy`` files (``numpy.lib.format.open_memmap``) can produce it
      frame block by frame block without ever holding the trajectory in memory;
    * ``<filename minus .npy>.top.npz``: topology arrays and, if present, the unit-cell vectors
      ``[n_frames, 3, 3]`` (36 bytes per frame).

    Returns the side-car file name."""
    filename = str(filename)
    if not filename.lower().endswith(".npy"):
        raise ValueError("save_npy writes a .npy coordinate file")
    out = np.lib.format.open_memmap(filename, mode="w+", dtype=np.float32, shape=tuple(trajectory.xyz.shape))
    step = max(1, (64 << 20) // max(1, trajectory.xyz.shape[1] * 12))
    for f0 in range(0, trajectory.xyz.shape[0], step):          # block copy: works for memory-mapped sources too
        out[f0:f0 + step] = trajectory.xyz[f0:f0 + step]
    out.flush()
    del out
    top = trajectory.topology
    arrays = {"atom_residue": top.atom_residue, "residue_chain": top.residue_chain,
              "atom_names": np.asarray(top.atom_names, dtype="U8"),
              "atom_elements": np.asarray(top.atom_elements, dtype="U4"),
              "residue_names": np.asarray(top.residue_names, dtype="U8"),
              "residue_resseq": np.asarray(top.residue_resseq, dtype=np.int64)}
    if trajectory.unitcell_vectors is not None:
        arrays["unitcell_vectors"] = trajectory.unitcell_vectors
    sidecar = filename[:-4] + ".top.npz"
    np.savez(sidecar, **arrays)
    return sidecar


def load_npy(filename, top=None):
    """Memory-mapped trajectory: ``xyz`` is a read-only view of the file, so slicing frames
    (``load(...)[subslice]``, the reference's ``load_trajectory_task``, frequency_task.py:70-88)
    touches no data, and the engine's staged host path (``cmb_accumulate_host`` /
    ``cmb_accumulate_host_gather``: host threads pack chunk k+1 into pinned staging while chunk k is
    copied and processed) pages in exactly the frames a rank works on.  A trajectory larger than
    host memory therefore streams through the GPU without ever being materialised.

    ``top`` (MDTraj's keyword): a :class:`Topology`, a trajectory object, or a ``.npz`` / ``.pdb``
    file; default: the side-car ``<name>.top.npz`` written by :func:`save_npy`."""
    filename = str(filename)
    xyz = np.load(filename, mmap_mode="r", allow_pickle=False)
    if xyz.dtype != np.float32 or xyz.ndim != 3 or xyz.shape[2] != 3 or not xyz.flags["C_CONTIGUOUS"]:
        raise IOError("%r must hold C-ordered float32 coordinates [n_frames, n_atoms, 3]" % (filename,))
    cell = None
    if top is None:
        top = filename[:-4] + ".top.npz"
    if isinstance(top, str):
        if top.lower().endswith(".npz"):
            data = np.load(top, allow_pickle=False)
            cell = data["unitcell_vectors"] if "unitcell_vectors" in data else None
            top = _topology_from_npz(data)
        else:
            top = load(top).topology
    elif hasattr(top, "topology"):
        top = top.topology
    return Trajectory(xyz, top, cell)


def load(filename, **kwargs):
    """Loader used by the chunked runner: ``.pdb`` text, ``.npz`` (xyz, unitcell_vectors,
    atom_residue, residue_chain[, atom_elements, residue_names]) or a memory-mapped ``.npy``
    coordinate file with its topology side-car (:func:`load_npy`, :func:`save_npy`)."""
    name = str(filename).lower()
    if name.endswith(".pdb"):
        return load_pdb(filename)
    if name.endswith(".npy"):
        return load_npy(filename, top=kwargs.get("top"))
    if name.endswith(".npz"):
        data = np.load(filename, allow_pickle=False)
        cell = data["unitcell_vectors"] if "unitcell_vectors" in data else None
        return Trajectory(data["xyz"], _topology_from_npz(data), cell)
    raise IOError("unsupported trajectory file %r (only .pdb, .npz and .npy)" % (filename,))


def topology_arrays(topology):
    """(atom -> residue index, atom -> chain index) for any mdtraj-like topology."""
    if isinstance(topology, Topology):
        return topology.atom_residue_chain()
    n = topology.n_atoms
    atom_res = np.empty(n, dtype=np.int32)
    atom_chain = np.empty(n, dtype=np.int32)
    for atom in topology.atoms:
        atom_res[atom.index] = atom.residue.index
        atom_chain[atom.index] = atom.residue.chain.index
    return atom_res, atom_chain
