"""The binding a maintainer of the REFERENCE package would add -- executable form of INTEGRATION.md B.

``bind(contact_map)`` takes the reference's own package (``import contact_map``) and returns
subclasses of ITS ``ContactFrequency`` / ``ContactTrajectory`` that override exactly the two private
hooks the reference's Dask runner overrides (/root/reference/contact_map/dask_runner.py:116-119,
192-202):

* ``ContactFrequency._build_contact_map(self, trajectory) -> (atom Counter, residue Counter)``
  (contact_map/contact_map.py:717-744)
* ``ContactTrajectory._build_contacts(self, trajectory) -> iterable[(atom Counter, residue Counter)]``
  (contact_map/contact_trajectory.py:129-130, 7-44)

Everything below the hooks goes through ``include/cmb200.h`` with ``ctypes`` and plain numpy
buffers -- no torch, no other module of this package besides the loader of the shared library
(``_lib``: ``ctypes.CDLL`` + argument types).  Device memory comes from ``cmb_device_alloc``.
Any base package with the reference's interface works (the tests bind this package's own classes,
whose hooks have the same contract, so that the binding is exercised on GPU boxes where the
reference is not installed; ``tests/test_reference_binding.py`` binds the real reference when it is
importable).

The result containers are the reference's: ``collections.Counter`` keyed by ``frozenset`` of REAL
atom / residue indices with integer frame counts, which the inherited ``ContactFrequency`` code
turns into ``ContactCount`` objects.
"""
import collections
import ctypes

import numpy as np

from . import _lib


class CmbError(RuntimeError):
    pass


class _Context(object):
    """One ``cmb_ctx`` per process and device (the C ABI's ownership rule)."""

    _instances = {}

    def __init__(self, device=0):
        self.L = _lib.lib()
        self.ctx = ctypes.c_void_p()
        rc = self.L.cmb_ctx_create(int(device), ctypes.byref(self.ctx))
        if rc != 0:
            raise CmbError("cmb_ctx_create failed (%d): %s" % (rc, (self.L.cmb_last_error(None) or b"").decode()))

    @classmethod
    def get(cls, device=0):
        if device not in cls._instances:
            cls._instances[device] = cls(device)
        return cls._instances[device]

    def check(self, rc, what):
        if rc != 0:
            raise CmbError("%s failed (%d): %s" % (what, rc, (self.L.cmb_last_error(self.ctx) or b"").decode()))

    # ---- device buffers ------------------------------------------------------------------
    def alloc(self, n_bytes):
        ptr = ctypes.c_void_p()
        self.check(self.L.cmb_device_alloc(self.ctx, int(n_bytes), ctypes.byref(ptr)), "cmb_device_alloc")
        return ptr

    def free(self, ptr):
        self.L.cmb_device_free(self.ctx, ptr)

    def to_host(self, d_ptr, dtype, count, offset_bytes=0):
        out = np.empty(int(count), dtype=dtype)
        if count:
            src = ctypes.c_void_p(d_ptr.value + int(offset_bytes))
            self.check(self.L.cmb_copy_to_host(self.ctx, out.ctypes.data_as(ctypes.c_void_p), src,
                                               out.nbytes, None), "cmb_copy_to_host")
        return out

    # ---- the system: ContactObject.__init__ (contact_map.py:120-150) -----------------------
    def set_system(self, contact_object):
        """Per-used-atom arrays of a reference ``ContactObject`` -> ``cmb_set_system``.
        Returns (used atoms, compact residue id -> real residue index)."""
        top = contact_object.topology
        used = np.asarray(sorted(set(contact_object.query) | set(contact_object.haystack)), dtype=np.int64)
        res = np.array([top.atom(int(a)).residue.index for a in used], dtype=np.int32)
        chain = np.array([top.atom(int(a)).residue.chain.index for a in used], dtype=np.int32)
        role = (np.isin(used, list(contact_object.query)).astype(np.uint8)
                | (np.isin(used, list(contact_object.haystack)).astype(np.uint8) << 1)).astype(np.uint8)
        self.check(self.L.cmb_set_system(self.ctx, len(used), res.ctypes.data_as(ctypes.c_void_p),
                                         chain.ctypes.data_as(ctypes.c_void_p),
                                         role.ctypes.data_as(ctypes.c_void_p), float(contact_object.cutoff),
                                         int(contact_object.n_neighbors_ignored)), "cmb_set_system")
        n_res = self.L.cmb_n_residues(self.ctx)
        rmap = np.empty(n_res, dtype=np.int32)
        self.check(self.L.cmb_residue_map(self.ctx, rmap.ctypes.data_as(ctypes.c_void_p)), "cmb_residue_map")
        return used, rmap.astype(np.int64)

    @staticmethod
    def host_trajectory(trajectory):
        """float32 C-contiguous coordinates (all atoms) and unit-cell vectors or None."""
        xyz = np.ascontiguousarray(trajectory.xyz, dtype=np.float32)
        have_cell = getattr(trajectory, "_have_unitcell", None)
        box = trajectory.unitcell_vectors if (have_cell is None or have_cell) else None
        if box is not None:
            box = np.ascontiguousarray(box, dtype=np.float32).reshape(-1, 3, 3)
        return xyz, box


def _pairs_to_counter(flat_index, counts, stride, index_map):
    lo, hi = np.divmod(np.asarray(flat_index, dtype=np.int64), stride)
    a, b = index_map[lo], index_map[hi]
    return collections.Counter({frozenset((int(i), int(j))): int(c) for i, j, c in zip(a, b, counts)})


def build_contact_map(contact_object, trajectory, device=0):
    """``_build_contact_map`` through the C ABI: cmb_set_system -> zeroed tables ->
    cmb_accumulate_host_gather (the fused ``_atom_slice``, atom_indexer.py:5-22) -> cmb_export_nonzero."""
    C = _Context.get(device)
    L = C.L
    used, rmap = C.set_system(contact_object)
    xyz, box = C.host_trajectory(trajectory)
    a_elems, r_elems = ctypes.c_int64(0), ctypes.c_int64(0)
    C.check(L.cmb_table_sizes(C.ctx, ctypes.byref(a_elems), ctypes.byref(r_elems)), "cmb_table_sizes")
    tables = C.alloc(4 * (a_elems.value + r_elems.value) + 16)
    atom_tab = tables
    res_tab = ctypes.c_void_p(tables.value + 4 * ((a_elems.value + 3) // 4 * 4))
    try:
        used32 = used.astype(np.int32)
        if len(xyz):
            C.check(L.cmb_accumulate_host_gather(
                C.ctx, xyz.ctypes.data_as(ctypes.c_void_p), int(xyz.shape[1]), used32.ctypes.data_as(ctypes.c_void_p),
                None if box is None else box.ctypes.data_as(ctypes.c_void_p), len(xyz), 0, 0, atom_tab, res_tab),
                "cmb_accumulate_host_gather")
        out = []
        for tab, n_elems, stride, index_map in ((atom_tab, a_elems.value, len(used), used),
                                                (res_tab, r_elems.value, len(rmap), rmap)):
            cap = max(16, min(n_elems, 64 * len(used) + 1024))
            while True:
                d_idx, d_cnt, d_n = C.alloc(8 * cap), C.alloc(4 * cap), C.alloc(8)
                try:
                    C.check(L.cmb_export_nonzero(C.ctx, tab, n_elems, cap, d_idx, d_cnt, d_n, None), "cmb_export_nonzero")
                    n = int(C.to_host(d_n, np.uint64, 1)[0])
                    if n <= cap:
                        idx, cnt = C.to_host(d_idx, np.int64, n), C.to_host(d_cnt, np.uint32, n)
                        break
                    cap = n
                finally:
                    for p in (d_idx, d_cnt, d_n):
                        C.free(p)
            out.append(_pairs_to_counter(idx, cnt, stride, index_map))
        return out[0], out[1]
    finally:
        C.free(tables)


def build_contacts(contact_object, trajectory, device=0):
    """``_build_contacts`` through the C ABI: cmb_frame_contacts_host -> cmb_sort_keys ->
    cmb_keys_to_csr; returns the reference's per-frame list of (atom Counter, residue Counter)."""
    C = _Context.get(device)
    L = C.L
    used, rmap = C.set_system(contact_object)
    xyz, box = C.host_trajectory(trajectory)
    n_frames = len(xyz)
    if n_frames == 0:
        return []
    used32, rmap32 = used.astype(np.int32), rmap.astype(np.int32)
    caps = [max(4096, 128 * n_frames), max(4096, 64 * n_frames)]
    while True:
        d_keys = [C.alloc(8 * caps[0]), C.alloc(8 * caps[1])]
        d_counts = C.alloc(32)
        try:
            C.check(L.cmb_frame_contacts_host(
                C.ctx, xyz.ctypes.data_as(ctypes.c_void_p), int(xyz.shape[1]), used32.ctypes.data_as(ctypes.c_void_p),
                None if box is None else box.ctypes.data_as(ctypes.c_void_p), n_frames, 0, 0, 0,
                d_keys[0], caps[0], d_keys[1], caps[1], d_counts), "cmb_frame_contacts_host")
            counts = C.to_host(d_counts, np.uint64, 2)
            if int(counts[0]) <= caps[0] and int(counts[1]) <= caps[1]:
                frames = []
                for keys, n, index_map in ((d_keys[0], int(counts[0]), used32), (d_keys[1], int(counts[1]), rmap32)):
                    if n > 1:
                        C.check(L.cmb_sort_keys(C.ctx, keys, n, None), "cmb_sort_keys")
                    d_map = C.alloc(4 * len(index_map))
                    d_ptr, d_i, d_j = C.alloc(8 * (n_frames + 1)), C.alloc(4 * max(n, 1)), C.alloc(4 * max(n, 1))
                    try:
                        C.check(L.cmb_copy_to_device(C.ctx, d_map, index_map.ctypes.data_as(ctypes.c_void_p),
                                                     index_map.nbytes, None), "cmb_copy_to_device")
                        C.check(L.cmb_keys_to_csr(C.ctx, keys, n, n_frames, 0, d_map, d_ptr, d_i, d_j, None),
                                "cmb_keys_to_csr")
                        ptr = C.to_host(d_ptr, np.int64, n_frames + 1)
                        i, j = C.to_host(d_i, np.int32, n), C.to_host(d_j, np.int32, n)
                    finally:
                        for p in (d_map, d_ptr, d_i, d_j):
                            C.free(p)
                    frames.append([collections.Counter({frozenset((int(a), int(b))): 1
                                                        for a, b in zip(i[ptr[f]:ptr[f + 1]], j[ptr[f]:ptr[f + 1]])})
                                   for f in range(n_frames)])
                return list(zip(frames[0], frames[1]))
            caps = [max(caps[0], int(counts[0])), max(caps[1], int(counts[1]))]
        finally:
            for p in d_keys + [d_counts]:
                C.free(p)


def bind(reference_package, device=0):
    """Subclasses of ``reference_package.ContactFrequency`` / ``ContactTrajectory`` whose hooks
    run on the GPU through the C ABI.  Returns ``(B200ContactFrequency, B200ContactTrajectory)``."""

    class B200ContactFrequency(reference_package.ContactFrequency):
        def _build_contact_map(self, trajectory):                   # contact_map.py:717
            return build_contact_map(self, trajectory, device)

    class B200ContactTrajectory(reference_package.ContactTrajectory):
        def _build_contacts(self, trajectory):                      # contact_trajectory.py:129
            return build_contacts(self, trajectory, device)

    return B200ContactFrequency, B200ContactTrajectory
