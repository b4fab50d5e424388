"""NearestAtoms / MinimumDistanceCounter on the GPU.

Mirror of /root/reference/contact_map/min_dist.py (NearestAtoms :5-106,
MinimumDistanceCounter :109-209).  The reference materialises ``[n_frames, n_q*n_h]``
float32 distances with ``md.compute_distances`` and reduces them with numpy; here a kernel
reduces every frame on the fly (``cmb_min_dist``) and one kernel finds every atom's nearest
allowed neighbour (``cmb_nearest``).
"""
import collections
import itertools

import numpy as np

from .contact_map import _split_trajectory
from .trajectory import topology_arrays

try:
    import torch
except ImportError:  # pragma: no cover
    torch = None


def _is_orthogonal(box):
    """np.allclose(unitcell_angles, 90), decided from the box vectors (MDTraj's python wrapper
    of compute_distances makes the orthogonal / triclinic choice this way)."""
    if box is None:
        return True
    box = np.asarray(box.cpu().numpy() if (torch is not None and isinstance(box, torch.Tensor)) else box,
                     dtype=np.float64).reshape(-1, 3, 3)
    a, b, c = box[:, 0], box[:, 1], box[:, 2]
    def angle(u, v):
        cosv = np.einsum("ij,ij->i", u, v) / (np.linalg.norm(u, axis=1) * np.linalg.norm(v, axis=1))
        return np.degrees(np.arccos(np.clip(cosv, -1.0, 1.0)))
    angles = np.stack([angle(b, c), angle(c, a), angle(a, b)], axis=1).astype(np.float32)
    return bool(np.allclose(angles, 90))


def _to_device(xyz, box, engine):
    if torch is not None and isinstance(xyz, torch.Tensor) and xyz.is_cuda:
        d_xyz = xyz.contiguous()
    else:
        d_xyz = torch.as_tensor(np.ascontiguousarray(xyz, dtype=np.float32), device=engine.device)
    d_box = None
    if box is not None:
        if torch is not None and isinstance(box, torch.Tensor) and box.is_cuda:
            d_box = box.reshape(-1, 3, 3).contiguous()
        else:
            d_box = torch.as_tensor(np.ascontiguousarray(box, dtype=np.float32).reshape(-1, 3, 3),
                                    device=engine.device)
    return d_xyz, d_box


class NearestAtoms(object):
    """Nearest atom (within a cutoff) to every atom of one frame (reference :5-106).

    ``nearest`` maps atom index -> index of its nearest allowed neighbour and
    ``nearest_distance`` maps atom index -> that distance (``np.float32``); atoms without an
    allowed neighbour inside the cutoff are absent.
    """

    def __init__(self, trajectory, cutoff, frame_number=0, excluded=None):
        self.cutoff = cutoff
        self.frame_number = frame_number
        self.excluded = self._parse_excluded(excluded, trajectory)
        self.nearest, self.nearest_distance = self._calculate_nearest(
            trajectory, self.cutoff, self.frame_number, self.excluded, _raw_excluded=excluded)

    @staticmethod
    def _parse_excluded(excluded, trajectory):
        n_atoms = trajectory.topology.n_atoms
        if excluded is not None and len(excluded) == 0:
            return {idx: [idx] for idx in range(n_atoms)}
        if excluded is None:
            atom_res, _ = topology_arrays(trajectory.topology)
            order = np.argsort(atom_res, kind="stable")
            groups = np.split(order, np.nonzero(np.diff(atom_res[order]))[0] + 1)
            excluded = {}
            for members in groups:
                members = [int(m) for m in members]
                for idx in members:
                    excluded[idx] = members
        return excluded

    @staticmethod
    def _calculate_nearest(trajectory, cutoff, frame_number, excluded, _raw_excluded="unknown"):
        from .engine import get_engine
        engine = get_engine()
        xyz, box = _split_trajectory(trajectory)
        n_atoms = xyz.shape[1]
        frame_box = None if box is None else box[frame_number:frame_number + 1] \
            if frame_number != -1 else box[-1:]
        d_xyz, d_box = _to_device(xyz[frame_number], frame_box, engine)
        # md.compute_distances(trajectory, ...)[frame_number] (min_dist.py:63-64): MDTraj decides
        # orthorhombic vs triclinic arithmetic from ALL frames of the trajectory it is given
        orthogonal = _is_orthogonal(box)
        if _raw_excluded is None:
            atom_res, _ = topology_arrays(trajectory.topology)
            mode, kwargs = 1, {"atom_res": atom_res}
        elif _raw_excluded is not None and not isinstance(_raw_excluded, str) and len(_raw_excluded) == 0:
            mode, kwargs = 0, {}
        else:
            ptr = np.zeros(n_atoms + 1, dtype=np.int64)
            chunks = []
            for atom in range(n_atoms):
                members = np.asarray(list(excluded[atom]), dtype=np.int32)   # KeyError like the reference
                chunks.append(members)
                ptr[atom + 1] = ptr[atom] + members.shape[0]
            idx = np.concatenate(chunks) if chunks else np.empty(0, np.int32)
            if idx.shape[0] == 0:
                idx = np.zeros(1, dtype=np.int32)
            mode, kwargs = 2, {"excl_ptr": ptr, "excl_idx": idx}
        near, dist = engine.nearest(d_xyz.reshape(n_atoms, 3), None if d_box is None else d_box.reshape(9),
                                    cutoff, orthogonal, mode, **kwargs)
        near, dist = near.cpu().numpy(), dist.cpu().numpy()
        found = np.nonzero(near >= 0)[0]
        nearest = {int(a): int(near[a]) for a in found}
        nearest_distance = {int(a): dist[a] for a in found}
        return nearest, nearest_distance

    @property
    def sorted_distances(self):
        """[(atom, nearest atom, distance)] sorted by distance."""
        listed = [(atom, self.nearest[atom], dist) for atom, dist in self.nearest_distance.items()]
        return sorted(listed, key=lambda tup: tup[2])


class MinimumDistanceCounter(object):
    """Which query/haystack atom pair is closest in each frame (reference :109-209)."""

    def __init__(self, trajectory, query, haystack):
        self.topology = trajectory.topology
        self._query = [int(q) for q in query]
        self._haystack = [int(h) for h in haystack]
        self._atom_pairs = None
        self.minimum_distances, self._min_pairs = self._compute_from_product(
            trajectory, self._query, self._haystack)

    @property
    def atom_pairs(self):
        """Query-major product of query and haystack (reference :138), built on demand."""
        if self._atom_pairs is None:
            self._atom_pairs = list(itertools.product(self._query, self._haystack))
        return self._atom_pairs

    @staticmethod
    def _compute_from_product(trajectory, query, haystack):
        from .engine import get_engine
        engine = get_engine()
        xyz, box = _split_trajectory(trajectory)
        n_atoms = xyz.shape[1]
        pairs = np.asarray(list(query) + list(haystack), dtype=np.int64)
        if pairs.size and (pairs.min() < 0 or pairs.max() >= n_atoms):
            raise ValueError("atom_pairs must be between 0 and %d" % n_atoms)
        d_xyz, d_box = _to_device(xyz, box, engine)
        dmin, darg = engine.min_dist(d_xyz, d_box, query, haystack, _is_orthogonal(box))
        return dmin.cpu().numpy(), darg.cpu().numpy()

    @staticmethod
    def _compute_from_atom_pairs(trajectory, atom_pairs):
        """The reference's seam for alternative constructors (min_dist.py:142-168): minimum
        distance and first arg-min index per frame over an arbitrary list of atom pairs, i.e.
        ``D.min(axis=1), D.argmin(axis=1)`` of ``D = md.compute_distances(trajectory,
        atom_pairs)`` -- returned as (minimum_distances, min_pairs) like the reference."""
        from .engine import get_engine
        engine = get_engine()
        xyz, box = _split_trajectory(trajectory)
        n_atoms = xyz.shape[1]
        pairs = np.asarray(list(atom_pairs) if not isinstance(atom_pairs, np.ndarray) else atom_pairs,
                           dtype=np.int64).reshape(-1, 2)
        if pairs.shape[0] == 0:
            raise ValueError("atom_pairs must not be empty")
        if pairs.min() < 0 or pairs.max() >= n_atoms:
            raise ValueError("atom_pairs must be between 0 and %d" % n_atoms)
        d_xyz, d_box = _to_device(xyz, box, engine)
        dmin, darg = engine.min_dist_pairs(d_xyz, d_box, pairs, _is_orthogonal(box))
        return dmin.cpu().numpy(), darg.cpu().numpy()

    def _remap(self, pair_number):
        n_h = len(self._haystack)
        q, h = self._query[int(pair_number) // n_h], self._haystack[int(pair_number) % n_h]
        return (self.topology.atom(q), self.topology.atom(h))

    @property
    def atom_history(self):
        """Closest (query atom, haystack atom) pair of every frame."""
        return [self._remap(k) for k in self._min_pairs]

    @property
    def atom_count(self):
        return collections.Counter(self.atom_history)

    @property
    def residue_history(self):
        return [(a[0].residue, a[1].residue) for a in self.atom_history]

    @property
    def residue_count(self):
        return collections.Counter(self.residue_history)
