"""Trajectory container + multi-MODEL PDB reader for the oracle's mdtraj stand-in.

Unit conventions follow MDTraj: coordinates and cell lengths in nanometres
(PDB Angstrom * 0.1 evaluated in float64, then narrowed to float32), angles in
degrees, ``unitcell_vectors`` as row vectors ``[F, 3, 3]``.
"""
import numpy as np

from . import element as elem
from .topology import Topology


def lengths_and_angles_to_box_vectors(a_length, b_length, c_length, alpha, beta, gamma):
    """MDTraj's published lengths/angles -> reduced-form box vectors recipe."""
    a_length = np.asarray(a_length, dtype=np.float64)
    b_length = np.asarray(b_length, dtype=np.float64)
    c_length = np.asarray(c_length, dtype=np.float64)
    alpha = np.asarray(alpha, dtype=np.float64) * np.pi / 180.0
    beta = np.asarray(beta, dtype=np.float64) * np.pi / 180.0
    gamma = np.asarray(gamma, dtype=np.float64) * np.pi / 180.0
    zeros = np.zeros_like(a_length)
    a = np.array([a_length, zeros, zeros])
    b = np.array([b_length * np.cos(gamma), b_length * np.sin(gamma), zeros])
    cx = c_length * np.cos(beta)
    cy = c_length * (np.cos(alpha) - np.cos(beta) * np.cos(gamma)) / np.sin(gamma)
    cz = np.sqrt(c_length * c_length - cx * cx - cy * cy)
    c = np.array([cx, cy, cz])
    tol = 1e-6
    for v in (a, b, c):
        v[np.logical_and(v > -tol, v < tol)] = 0.0
    return a.T, b.T, c.T


def box_vectors_to_lengths_and_angles(a, b, c):
    a_length = np.sqrt(np.sum(a * a, axis=-1))
    b_length = np.sqrt(np.sum(b * b, axis=-1))
    c_length = np.sqrt(np.sum(c * c, axis=-1))
    alpha = np.arccos(np.einsum("...i,...i", b, c) / (b_length * c_length)) * 180.0 / np.pi
    beta = np.arccos(np.einsum("...i,...i", c, a) / (c_length * a_length)) * 180.0 / np.pi
    gamma = np.arccos(np.einsum("...i,...i", a, b) / (a_length * b_length)) * 180.0 / np.pi
    return a_length, b_length, c_length, alpha, beta, gamma


class Trajectory(object):
    def __init__(self, xyz, topology, time=None, unitcell_lengths=None, unitcell_angles=None):
        xyz = np.asarray(xyz, dtype=np.float32)
        if xyz.ndim == 2:
            xyz = xyz[np.newaxis]
        self._xyz = np.ascontiguousarray(xyz)
        self._topology = topology
        if time is None:
            time = np.arange(self._xyz.shape[0], dtype=np.float32)
        self._time = np.asarray(time, dtype=np.float32).reshape(-1)
        self._unitcell_lengths = None
        self._unitcell_angles = None
        if unitcell_lengths is not None:
            self._unitcell_lengths = np.asarray(unitcell_lengths, dtype=np.float32).reshape(-1, 3)
            self._unitcell_angles = np.asarray(unitcell_angles, dtype=np.float32).reshape(-1, 3)
        if topology is not None and topology.n_atoms != self._xyz.shape[1]:
            raise ValueError("topology has %d atoms but xyz has %d"
                             % (topology.n_atoms, self._xyz.shape[1]))

    # -- basic properties ----------------------------------------------
    @property
    def xyz(self):
        return self._xyz

    @xyz.setter
    def xyz(self, value):
        self._xyz = np.ascontiguousarray(np.asarray(value, dtype=np.float32))

    @property
    def topology(self):
        return self._topology

    @topology.setter
    def topology(self, value):
        self._topology = value

    top = topology

    @property
    def time(self):
        return self._time

    @property
    def n_frames(self):
        return self._xyz.shape[0]

    @property
    def n_atoms(self):
        return self._xyz.shape[1]

    @property
    def n_residues(self):
        return self._topology.n_residues

    def __len__(self):
        return self.n_frames

    # -- unit cell -----------------------------------------------------
    @property
    def _have_unitcell(self):
        return self._unitcell_lengths is not None and self._unitcell_angles is not None

    @property
    def unitcell_lengths(self):
        return self._unitcell_lengths

    @property
    def unitcell_angles(self):
        return self._unitcell_angles

    @property
    def unitcell_vectors(self):
        if not self._have_unitcell:
            return None
        L, A = self._unitcell_lengths, self._unitcell_angles
        v1, v2, v3 = lengths_and_angles_to_box_vectors(L[:, 0], L[:, 1], L[:, 2],
                                                       A[:, 0], A[:, 1], A[:, 2])
        return np.swapaxes(np.dstack((v1, v2, v3)), 1, 2)

    @unitcell_vectors.setter
    def unitcell_vectors(self, vectors):
        if vectors is None:
            self._unitcell_lengths = None
            self._unitcell_angles = None
            return
        vectors = np.asarray(vectors, dtype=np.float64).reshape(-1, 3, 3)
        a, b, c, al, be, ga = box_vectors_to_lengths_and_angles(vectors[:, 0], vectors[:, 1],
                                                                vectors[:, 2])
        self._unitcell_lengths = np.vstack((a, b, c)).T.astype(np.float32)
        self._unitcell_angles = np.vstack((al, be, ga)).T.astype(np.float32)

    # -- slicing -------------------------------------------------------
    def __getitem__(self, key):
        return self.slice(key)

    def slice(self, key, copy=True):
        if isinstance(key, (int, np.integer)):
            key = [int(key)]
        xyz = self._xyz[key]
        time = self._time[key]
        lengths = angles = None
        if self._have_unitcell:
            lengths = self._unitcell_lengths[key]
            angles = self._unitcell_angles[key]
        if copy:
            xyz = xyz.copy()
            time = time.copy()
            topology = self._topology.copy()
            if lengths is not None:
                lengths, angles = lengths.copy(), angles.copy()
        else:
            topology = self._topology
        return Trajectory(xyz, topology, time=time, unitcell_lengths=lengths,
                          unitcell_angles=angles)

    def atom_slice(self, atom_indices, inplace=False):
        idx = np.asarray(list(atom_indices), dtype=int)
        xyz = np.array(self._xyz[:, idx], order="C")
        topology = self._topology.subset(idx)
        lengths = angles = None
        if self._have_unitcell:
            lengths, angles = self._unitcell_lengths.copy(), self._unitcell_angles.copy()
        return Trajectory(xyz, topology, time=self._time.copy(), unitcell_lengths=lengths,
                          unitcell_angles=angles)

    def join(self, other, check_topology=True, discard_overlapping_frames=False):
        others = other if isinstance(other, (list, tuple)) else [other]
        return join([self] + list(others), check_topology=check_topology)

    def __add__(self, other):
        return self.join(other)

    def __eq__(self, other):
        if not isinstance(other, Trajectory):
            return False
        if self._topology != other._topology:
            return False
        if self._xyz.shape != other._xyz.shape or not np.array_equal(self._xyz, other._xyz):
            return False
        if self._have_unitcell != other._have_unitcell:
            return False
        if self._have_unitcell and not (
                np.array_equal(self._unitcell_lengths, other._unitcell_lengths)
                and np.array_equal(self._unitcell_angles, other._unitcell_angles)):
            return False
        return True

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = None

    def __repr__(self):
        return "<oracle.mdtraj.Trajectory with %d frames, %d atoms>" % (self.n_frames, self.n_atoms)


def join(trajs, check_topology=True, discard_overlapping_frames=False):
    trajs = list(trajs)
    first = trajs[0]
    if check_topology:
        for t in trajs[1:]:
            if t.topology != first.topology:
                raise ValueError("topologies do not match")
    xyz = np.concatenate([t.xyz for t in trajs], axis=0)
    time = np.concatenate([t.time for t in trajs], axis=0)
    lengths = angles = None
    if all(t._have_unitcell for t in trajs):
        lengths = np.concatenate([t.unitcell_lengths for t in trajs], axis=0)
        angles = np.concatenate([t.unitcell_angles for t in trajs], axis=0)
    return Trajectory(xyz, first.topology.copy(), time=time, unitcell_lengths=lengths,
                      unitcell_angles=angles)


# ---------------------------------------------------------------------
# PDB reader (fixed-column ATOM/HETATM, CRYST1, MODEL/ENDMDL)
# ---------------------------------------------------------------------
def load_pdb(filename, stride=None, atom_indices=None, frame=None, **kwargs):
    topology = Topology()
    frames = []
    cells = []
    current = []
    cell = None
    first_model_done = False
    chain_by_id = {}
    last_res_key = None
    residue = None
    have_models = False

    def finish_model():
        nonlocal current, first_model_done
        if current:
            frames.append(current)
            cells.append(cell)
            first_model_done = True
        current = []

    with open(filename) as fh:
        for line in fh:
            rec = line[:6]
            if rec == "CRYST1":
                cell = (float(line[6:15]), float(line[15:24]), float(line[24:33]),
                        float(line[33:40]), float(line[40:47]), float(line[47:54]))
            elif rec.startswith("MODEL"):
                have_models = True
                finish_model()
            elif rec == "ENDMDL":
                finish_model()
            elif rec in ("ATOM  ", "HETATM"):
                x, y, z = float(line[30:38]), float(line[38:46]), float(line[46:54])
                current.append((x, y, z))
                if not first_model_done:
                    serial = int(line[6:11])
                    name = line[12:16].strip()
                    res_name = line[17:20].strip()
                    chain_id = line[21]
                    res_seq = int(line[22:26])
                    seg_id = line[72:76].strip() if len(line) > 72 else ""
                    symbol = line[76:78].strip() if len(line) > 76 else ""
                    if chain_id not in chain_by_id:
                        chain_by_id[chain_id] = topology.add_chain(chain_id)
                        last_res_key = None
                    chain = chain_by_id[chain_id]
                    key = (chain_id, res_seq, res_name)
                    if key != last_res_key:
                        residue = topology.add_residue(res_name, chain, res_seq, seg_id)
                        last_res_key = key
                    if symbol:
                        try:
                            element = elem.get_by_symbol(symbol)
                        except KeyError:
                            element = elem.virtual
                    else:
                        element = elem.get_by_symbol(name[0]) if name[0].upper() in "HCNOSPF" \
                            else elem.virtual
                    topology.add_atom(name, element, residue, serial=serial)
            elif rec.startswith("TER"):
                last_res_key = None
    finish_model()
    del have_models

    coords = np.array(frames, dtype=np.float64) * 0.1
    xyz = coords.astype(np.float32)
    lengths = angles = None
    if all(c is not None for c in cells) and cells:
        arr = np.array(cells, dtype=np.float64)
        lengths = (arr[:, :3] * 0.1).astype(np.float32)
        angles = arr[:, 3:].astype(np.float32)
    traj = Trajectory(xyz, topology, unitcell_lengths=lengths, unitcell_angles=angles)
    if frame is not None:
        traj = traj[frame]
    if stride is not None:
        traj = traj[::stride]
    if atom_indices is not None:
        traj = traj.atom_slice(atom_indices)
    return traj


def load(filename, **kwargs):
    if str(filename).lower().endswith(".pdb"):
        return load_pdb(filename, **kwargs)
    raise IOError("the oracle shim only reads .pdb files (got %r)" % (filename,))
