"""ContactTrajectory: per-frame contacts, stored as CSR arrays filled by the GPU.

Mirror of /root/reference/contact_map/contact_trajectory.py (``_build_contacts`` :7-44,
``ContactTrajectory`` :47-218, ``MutableContactTrajectory`` :221-252, ``WindowedIterator``
:255-329, ``RollingContactFrequency`` :332-392).

The reference builds one Python ``ContactFrequency`` (two ``Counter`` objects plus a full
``ContactObject.__init__``) per frame.  Here the engine emits 64-bit (frame, lo, hi) keys,
one device radix sort orders them, and frames are materialised as 1-frame
``ContactFrequency`` objects only when indexed.
"""
import collections
from collections import abc

import numpy as np

from .contact_count import PairTable
from .contact_map import (ContactFrequency, ContactObject, _engine_system, _split_trajectory,
                          _used_coordinates)

try:
    import torch
except ImportError:  # pragma: no cover
    torch = None



# layout of the 64-bit per-frame keys (cmb_frame_contacts): frame << 40 | lo << 20 | hi
KEY_MAX_FRAMES = 1 << 24
KEY_MAX_INDEX = 1 << 20


class FrameCSR(object):
    """Per-frame unordered pairs: ``ptr[f]:ptr[f+1]`` indexes ``i``/``j`` (i < j, real indices)."""

    def __init__(self, n_frames, frame, i, j):
        self.n_frames = int(n_frames)
        self.ptr = np.searchsorted(frame, np.arange(self.n_frames + 1), side="left")
        self.i = np.asarray(i, dtype=np.int64)
        self.j = np.asarray(j, dtype=np.int64)

    @classmethod
    def from_arrays(cls, n_frames, ptr, i, j):
        """Ready-made CSR arrays (``cmb_keys_to_csr``): nothing is recomputed on the host."""
        self = cls.__new__(cls)
        self.n_frames = int(n_frames)
        self.ptr, self.i, self.j = ptr, i, j
        return self

    @classmethod
    def from_keys(cls, n_frames, keys, index_map):
        k = np.asarray(keys, dtype=np.uint64)                      # unsigned shifts: frame ids use bits 40..63
        index_map = np.asarray(index_map, dtype=np.int64)
        if int(n_frames) > KEY_MAX_FRAMES:
            raise ValueError("per-frame keys hold 24-bit frame ids: at most %d frames per key stream"
                             % KEY_MAX_FRAMES)
        if index_map.shape[0] > KEY_MAX_INDEX:
            raise ValueError("per-frame keys hold 20-bit indices: at most %d used atoms / residues" % KEY_MAX_INDEX)
        frame = (k >> np.uint64(40)).astype(np.int64)
        lo = index_map[((k >> np.uint64(20)) & np.uint64(0xFFFFF)).astype(np.int64)]
        hi = index_map[(k & np.uint64(0xFFFFF)).astype(np.int64)]
        # keys hold lo < hi in compact numbering; an increasing map (sorted used atoms) keeps that
        if index_map.shape[0] > 1 and not bool(np.all(index_map[1:] > index_map[:-1])):
            lo, hi = np.minimum(lo, hi), np.maximum(lo, hi)
        return cls(n_frames, frame, lo, hi)

    def frame_table(self, f):
        a, b = self.ptr[f], self.ptr[f + 1]
        return PairTable(self.i[a:b], self.j[a:b], np.ones(b - a, dtype=np.int64))

    def total_table(self, frames=None):
        """Sum over frames (all, or a slice) as integer counts."""
        if frames is None:
            a, b = 0, self.ptr[-1]
        else:
            a, b = self.ptr[frames.start], self.ptr[frames.stop]
        if b == a:
            return PairTable.empty()
        base = int(self.j[a:b].max()) + 1
        uniq, counts = np.unique(self.i[a:b].astype(np.int64) * base + self.j[a:b], return_counts=True)
        return PairTable(uniq // base, uniq % base, counts.astype(np.int64))


def _build_contacts(contact_object, trajectory):
    """GPU replacement of the reference's ``_build_contacts`` (contact_trajectory.py:7-44).

    Returns (atom FrameCSR, residue FrameCSR) in REAL atom / residue indices.  Host trajectories
    stream through the engine's pinned double-buffered pipeline (``cmb_frame_contacts_host``; the
    used atoms of a trajectory that still holds all atoms are gathered on the way), the keys are
    sorted and turned into CSR arrays on the device, and only those arrays come back.
    """
    from .engine import get_engine
    engine = get_engine()
    used = _engine_system(contact_object, engine)
    xyz, box = _split_trajectory(trajectory)
    xyz, box, on_device = _used_coordinates(xyz, box, used, contact_object.topology.n_atoms, engine,
                                            keep_full_host=True)
    n_frames = xyz.shape[0]
    if n_frames >= KEY_MAX_FRAMES:
        raise ValueError("ContactTrajectory handles fewer than %d frames per call (24-bit frame ids in "
                         "the per-frame keys); split the trajectory and join the parts" % KEY_MAX_FRAMES)
    if used.shape[0] > KEY_MAX_INDEX:
        raise ValueError("per-frame keys hold 20-bit indices: at most %d used atoms" % KEY_MAX_INDEX)
    res_map = engine.residue_map
    if n_frames == 0:
        empty = (np.zeros(1, np.int64), np.empty(0, np.int32), np.empty(0, np.int32))
        return FrameCSR.from_arrays(0, *empty), FrameCSR.from_arrays(0, *empty)
    gather = (not on_device) and xyz.shape[1] != used.shape[0]
    atoms, residues = engine.frame_csr(xyz, box, used, res_map, used=used if gather else None)
    return FrameCSR.from_arrays(n_frames, *atoms), FrameCSR.from_arrays(n_frames, *residues)


class _LazyFrames(abc.Sequence):
    """Sequence of 1-frame ContactFrequency objects built on first access."""

    def __init__(self, owner, atoms, residues):
        # (the parameters are copied, not the owner: a reference back to the ContactTrajectory would
        # be a cycle, and the CSR arrays -- pinned host memory -- would wait for the cycle collector)
        self._params = dict(topology=owner.topology, query=owner.query, haystack=owner.haystack,
                            cutoff=owner.cutoff, n_neighbors_ignored=owner.n_neighbors_ignored,
                            indexer=owner.indexer)
        self._atoms = atoms
        self._residues = residues
        self._cache = {}

    def __len__(self):
        return self._atoms.n_frames

    def _make(self, f):
        if f not in self._cache:
            self._cache[f] = ContactFrequency.from_contacts(
                self._atoms.frame_table(f), self._residues.frame_table(f), n_frames=1, **self._params)
        return self._cache[f]

    def __getitem__(self, num):
        if isinstance(num, slice):
            return [self._make(f) for f in range(*num.indices(len(self)))]
        if num < 0:
            num += len(self)
        if not 0 <= num < len(self):
            raise IndexError(num)
        return self._make(num)


class ContactTrajectory(ContactObject, abc.Sequence):
    """Track all the contacts over a trajectory, frame by frame (reference :47-218)."""
    _class_use_atom_slice = None

    def __init__(self, trajectory, query=None, haystack=None, cutoff=0.45, n_neighbors_ignored=2):
        super(ContactTrajectory, self).__init__(trajectory.topology, query, haystack, cutoff,
                                                n_neighbors_ignored)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", PendingDeprecationWarning)
            contacts = self._build_contacts(trajectory)
        self._install(contacts)

    def _install(self, contacts):
        if isinstance(contacts, tuple) and len(contacts) == 2 and isinstance(contacts[0], FrameCSR):
            self._csr = contacts
            self._contact_maps = _LazyFrames(self, contacts[0], contacts[1])
        else:   # iterable of (atom counter, residue counter) per frame (the reference's hook type)
            self._csr = None
            self._contact_maps = [
                ContactFrequency.from_contacts(atoms, residues, n_frames=1, topology=self.topology,
                                               query=self.query, haystack=self.haystack,
                                               cutoff=self.cutoff,
                                               n_neighbors_ignored=self.n_neighbors_ignored,
                                               indexer=self.indexer)
                for atoms, residues in contacts]

    def _build_contacts(self, trajectory):
        return _build_contacts(self, trajectory)

    def __getitem__(self, num):
        return self._contact_maps[num]

    def __len__(self):
        return len(self._contact_maps)

    def __hash__(self):
        return hash((super(ContactTrajectory, self).__hash__(),
                     tuple(frozenset(frame.counter.items()) for frame in self.atom_contacts),
                     tuple(frozenset(frame.counter.items()) for frame in self.residue_contacts)))

    def __eq__(self, other):
        return hash(self) == hash(other)

    def __ne__(self, other):
        return not self.__eq__(other)

    @classmethod
    def from_contacts(cls, atom_contacts, residue_contacts, topology, query=None, haystack=None,
                      cutoff=0.45, n_neighbors_ignored=2):
        maps = [ContactFrequency.from_contacts(a, r, n_frames=1, topology=topology, query=query,
                                               haystack=haystack, cutoff=cutoff,
                                               n_neighbors_ignored=n_neighbors_ignored)
                for a, r in zip(atom_contacts, residue_contacts)]
        return cls.from_contact_maps(maps)

    @classmethod
    def from_contact_maps(cls, maps):
        obj = cls.__new__(cls)
        ContactObject.__init__(obj, maps[0].topology, maps[0].query, maps[0].haystack,
                               maps[0].cutoff, maps[0].n_neighbors_ignored)
        for cmap in maps:
            obj._check_compatibility(cmap)
        obj._csr = None
        obj._contact_maps = list(maps)
        return obj

    @classmethod
    def join(cls, others):
        """Concatenate ContactTrajectory instances (reference :185-200)."""
        maps = []
        for other in others:
            maps.extend(list(other._contact_maps))
        return cls.from_contact_maps(maps)

    def to_dict(self):
        """Reference contact_trajectory.py:152-155."""
        return {"contact_maps": [cmap.to_dict() for cmap in self._contact_maps]}

    @classmethod
    def from_dict(cls, dct):
        """Reference contact_trajectory.py:157-162."""
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", PendingDeprecationWarning)
            return cls.from_contact_maps([ContactFrequency.from_dict(cmap) for cmap in dct["contact_maps"]])

    def to_json(self):
        import json
        return json.dumps(self.to_dict())

    @classmethod
    def from_json(cls, json_string):
        import json
        return cls.from_dict(json.loads(json_string))

    def contact_frequency(self):
        """ContactFrequency summed over all frames (reference :132-150); one reduce-by-key
        over the CSR arrays instead of per-frame ``add_contact_frequency`` calls."""
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", PendingDeprecationWarning)
            if self._csr is not None:
                return ContactFrequency.from_contacts(
                    self._csr[0].total_table(), self._csr[1].total_table(), n_frames=len(self),
                    topology=self.topology, query=self.query, haystack=self.haystack,
                    cutoff=self.cutoff, n_neighbors_ignored=self.n_neighbors_ignored)
            freq = ContactFrequency.from_contacts(
                collections.Counter(), collections.Counter(), n_frames=0, topology=self.topology,
                query=self.query, haystack=self.haystack, cutoff=self.cutoff,
                n_neighbors_ignored=self.n_neighbors_ignored)
            for cmap in self._contact_maps:
                freq.add_contact_frequency(cmap)
            return freq

    @property
    def atom_contacts(self):
        return [cmap.atom_contacts for cmap in self._contact_maps]

    @property
    def residue_contacts(self):
        return [cmap.residue_contacts for cmap in self._contact_maps]

    def rolling_frequency(self, window_size=1, step=1):
        return RollingContactFrequency(self, width=window_size, step=step)


class MutableContactTrajectory(ContactTrajectory, abc.MutableSequence):
    """Mutable version of ContactTrajectory (reference :221-252)."""

    def _install(self, contacts):
        super(MutableContactTrajectory, self)._install(contacts)
        self._contact_maps = list(self._contact_maps)
        self._csr = None

    def __setitem__(self, key, value):
        self._contact_maps[key] = value

    def __delitem__(self, key):
        del self._contact_maps[key]

    def insert(self, key, value):
        self._contact_maps.insert(key, value)

    def __hash__(self):
        return id(self)


class WindowedIterator(abc.Iterator):
    """(to_add, to_sub) slices of successive windows (reference :255-329)."""

    def __init__(self, length, width, step, slow_build):
        self.length = length
        self.width = width
        self.step = step
        self.slow_build = slow_build
        self.min = -1
        self.max = -1

    def __next__(self):
        if self.max + self.step >= self.length:
            raise StopIteration
        low = max(0, self.min)
        new_max = self.max + self.step
        if not self.slow_build:
            new_max = max(new_max, self.width - 1)
        new_min = max(low, new_max - self.width + 1)
        to_sub = slice(low, new_min)
        to_add = slice(self.max + 1, new_max + 1)
        self.min, self.max = new_min, new_max
        return to_add, to_sub


class RollingContactFrequency(abc.Iterator):
    """Rolling-window contact frequencies over a ContactTrajectory (reference :332-392)."""
    _slow_build_iter = False

    def __init__(self, contact_trajectory, width=1, step=1):
        self.trajectory = contact_trajectory
        self.width = width
        self.step = step
        self.slow_build_iter = self._slow_build_iter
        self._window_iter = None
        self._contact_map = None

    def __iter__(self):
        import warnings
        self._window_iter = WindowedIterator(length=len(self.trajectory), width=self.width,
                                             step=self.step, slow_build=self.slow_build_iter)
        t = self.trajectory
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", PendingDeprecationWarning)
            self._contact_map = ContactFrequency.from_contacts(
                collections.Counter(), collections.Counter(), topology=t.topology, query=t.query,
                haystack=t.haystack, cutoff=t.cutoff, n_neighbors_ignored=t.n_neighbors_ignored,
                n_frames=0)
        return self

    def __next__(self):
        import warnings
        to_add, to_sub = next(self._window_iter)
        t, cmap = self.trajectory, self._contact_map
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", PendingDeprecationWarning)
            csr = getattr(t, "_csr", None)
            if csr is not None:
                # window delta straight from the CSR arrays
                for frames, sign in ((to_add, +1), (to_sub, -1)):
                    lo, hi, _ = frames.indices(len(t))
                    if hi > lo:
                        delta = ContactFrequency.from_contacts(
                            csr[0].total_table(slice(lo, hi)), csr[1].total_table(slice(lo, hi)),
                            n_frames=hi - lo, topology=t.topology, query=t.query,
                            haystack=t.haystack, cutoff=t.cutoff,
                            n_neighbors_ignored=t.n_neighbors_ignored)
                        cmap._combine(delta, sign)
            else:
                for frame in t[to_add]:
                    cmap.add_contact_frequency(frame)
                for frame in t[to_sub]:
                    cmap.subtract_contact_frequency(frame)
            return ContactFrequency.from_contacts(
                cmap._table_of("atom"), cmap._table_of("residue"), topology=cmap.topology,
                query=cmap.query, haystack=cmap.haystack, cutoff=cmap.cutoff,
                n_neighbors_ignored=cmap.n_neighbors_ignored, n_frames=cmap.n_frames)
