"""ctypes binding of ``oracle/cm_oracle.c`` (CPU ORACLE -- test infrastructure only).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module; the product package
``contact_map_b200`` never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcm_oracle.so")
_SRC = os.path.join(_HERE, "cm_oracle.c")
_LIB = None

_BASE_FLAGS = ["-O2", "-ffp-contract=off", "-mno-fma", "-fno-fast-math", "-shared", "-fPIC",
               "-std=gnu11"]


def build(force=False):
    """Compile the oracle with every float op rounded separately (-ffp-contract=off)."""
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(_SRC):
        return _SO
    compilers = [c for c in ("/usr/bin/gcc", "gcc", "cc") if c]
    last = None
    for cc in compilers:
        for omp in (["-fopenmp"], []):
            tmp = _SO + ".tmp%d" % os.getpid()        # atomic: concurrent builders never see a partial file
            cmd = [cc] + _BASE_FLAGS + omp + ["-o", tmp, _SRC, "-lm"]
            try:
                subprocess.run(cmd, check=True, capture_output=True, text=True)
                os.replace(tmp, _SO)
                return _SO
            except (OSError, subprocess.CalledProcessError) as exc:   # try the next recipe
                last = exc
    raise RuntimeError("could not build the oracle library: %r" % (last,))


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    build()
    L = ctypes.CDLL(_SO)
    c_f = ctypes.POINTER(ctypes.c_float)
    c_i32 = ctypes.POINTER(ctypes.c_int32)
    c_i64 = ctypes.POINTER(ctypes.c_int64)
    c_u8 = ctypes.POINTER(ctypes.c_uint8)
    c_u32 = ctypes.POINTER(ctypes.c_uint32)
    L.orc_version.restype = ctypes.c_int
    L.orc_max_threads.restype = ctypes.c_int
    L.orc_set_threads.restype = None
    L.orc_set_threads.argtypes = [ctypes.c_int]
    L.orc_nl_d2.restype = ctypes.c_float
    L.orc_nl_d2.argtypes = [c_f, c_f, c_f]
    L.orc_neighborlist_csr.restype = ctypes.c_int64
    L.orc_neighborlist_csr.argtypes = [c_f, ctypes.c_int, c_f, ctypes.c_double, c_i64, c_i32]
    L.orc_neighborlist_csr_cells.restype = ctypes.c_int64
    L.orc_neighborlist_csr_cells.argtypes = [c_f, ctypes.c_int, c_f, ctypes.c_double, c_i64, c_i32]
    sig = [c_f, ctypes.c_int, c_f, ctypes.c_double, c_i32, c_i32, c_u8, ctypes.c_int,
           ctypes.c_int64, c_i32, c_i32]
    L.orc_contacts_frame.restype = ctypes.c_int64
    L.orc_contacts_frame.argtypes = sig
    L.orc_contacts_frame_cells.restype = ctypes.c_int64
    L.orc_contacts_frame_cells.argtypes = sig
    L.orc_accumulate_dense.restype = None
    L.orc_accumulate_dense.argtypes = [c_f, c_f, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                       c_i32, c_i32, c_u8, ctypes.c_int, ctypes.c_int,
                                       c_u32, c_u32]
    L.orc_boundary_pairs.restype = ctypes.c_int64
    L.orc_boundary_pairs.argtypes = [c_f, ctypes.c_int, c_f, ctypes.c_double, c_i32, c_i32, c_u8,
                                     ctypes.c_int, ctypes.c_int, ctypes.c_int64, c_i32, c_i32, c_f]
    L.orc_dist_pair.restype = ctypes.c_float
    L.orc_dist_pair.argtypes = [c_f, c_f, c_f, ctypes.c_int]
    L.orc_distances.restype = None
    L.orc_distances.argtypes = [c_f, ctypes.c_int, ctypes.c_int, c_i32, ctypes.c_int64, c_f,
                                ctypes.c_int, c_f]
    L.orc_min_dist.restype = None
    L.orc_min_dist.argtypes = [c_f, ctypes.c_int, ctypes.c_int, c_i32, ctypes.c_int, c_i32,
                               ctypes.c_int, c_f, ctypes.c_int, c_f, c_i64]
    _LIB = L
    return L


def _p(arr, ctype):
    if arr is None:
        return None
    return arr.ctypes.data_as(ctypes.POINTER(ctype))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _box9(box):
    if box is None:
        return None
    b = np.ascontiguousarray(box, dtype=np.float32).reshape(-1)
    assert b.size % 9 == 0
    return b


# ---------------------------------------------------------------------
# thin numpy-level wrappers
# ---------------------------------------------------------------------
def nl_d2(pi, pj, box=None):
    pi, pj, b = _f32(pi), _f32(pj), _box9(box)
    return float(lib().orc_nl_d2(_p(pi, ctypes.c_float), _p(pj, ctypes.c_float),
                                 _p(b, ctypes.c_float)))


def neighborlist(xyz_frame, box, cutoff, cells=True):
    """MDTraj-style neighbour list of one frame: list of int arrays (one per atom).
    cells=True prunes candidates with a cell grid (identical result, see the tests)."""
    x = _f32(xyz_frame)
    n = x.shape[0]
    b = _box9(box)
    ptr = np.zeros(n + 1, dtype=np.int64)
    fn = lib().orc_neighborlist_csr_cells if cells else lib().orc_neighborlist_csr
    total = fn(_p(x, ctypes.c_float), n, _p(b, ctypes.c_float), float(cutoff),
               _p(ptr, ctypes.c_int64), None)
    idx = np.empty(max(int(total), 1), dtype=np.int32)
    fn(_p(x, ctypes.c_float), n, _p(b, ctypes.c_float), float(cutoff),
       _p(ptr, ctypes.c_int64), _p(idx, ctypes.c_int32))
    idx64 = idx.astype(np.int64)
    return [idx64[ptr[i]:ptr[i + 1]] for i in range(n)]


def contacts_frame(xyz_frame, box, cutoff, res, chain, role, n_ignored, cells=False):
    """Closed-form contacts of one frame; returns int32 array [n_pairs, 2], a < b, sorted."""
    x = _f32(xyz_frame)
    n = x.shape[0]
    b = _box9(box)
    res, chain = _i32(res), _i32(chain)
    role = np.ascontiguousarray(role, dtype=np.uint8)
    fn = lib().orc_contacts_frame_cells if cells else lib().orc_contacts_frame
    args = (_p(x, ctypes.c_float), n, _p(b, ctypes.c_float), float(cutoff),
            _p(res, ctypes.c_int32), _p(chain, ctypes.c_int32), _p(role, ctypes.c_uint8),
            int(n_ignored))
    total = fn(*args, 0, None, None)
    if total < 0:
        raise RuntimeError("oracle: singular box")
    oa = np.empty(max(int(total), 1), dtype=np.int32)
    ob = np.empty(max(int(total), 1), dtype=np.int32)
    fn(*args, int(total), _p(oa, ctypes.c_int32), _p(ob, ctypes.c_int32))
    return np.stack([oa[:total], ob[:total]], axis=1)


def accumulate_dense(xyz, box, cutoff, res, chain, role, n_ignored, n_res):
    """Dense uint32 tables over F frames: atoms [n, n] (a<b), residues [R, R] (min,max)."""
    x = _f32(xyz)
    F, n = x.shape[0], x.shape[1]
    b = _box9(box)
    res, chain = _i32(res), _i32(chain)
    role = np.ascontiguousarray(role, dtype=np.uint8)
    atom = np.zeros((n, n), dtype=np.uint32)
    resc = np.zeros((n_res, n_res), dtype=np.uint32)
    lib().orc_accumulate_dense(_p(x, ctypes.c_float), _p(b, ctypes.c_float), F, n, float(cutoff),
                               _p(res, ctypes.c_int32), _p(chain, ctypes.c_int32),
                               _p(role, ctypes.c_uint8), int(n_ignored), int(n_res),
                               _p(atom, ctypes.c_uint32), _p(resc, ctypes.c_uint32))
    return atom, resc


def boundary_pairs(xyz_frame, box, cutoff, res, chain, role, n_ignored, ulps=1, cap=1 << 20):
    x = _f32(xyz_frame)
    n = x.shape[0]
    b = _box9(box)
    res, chain = _i32(res), _i32(chain)
    role = np.ascontiguousarray(role, dtype=np.uint8)
    oa = np.empty(cap, dtype=np.int32)
    ob = np.empty(cap, dtype=np.int32)
    od = np.empty(cap, dtype=np.float32)
    k = lib().orc_boundary_pairs(_p(x, ctypes.c_float), n, _p(b, ctypes.c_float), float(cutoff),
                                 _p(res, ctypes.c_int32), _p(chain, ctypes.c_int32),
                                 _p(role, ctypes.c_uint8), int(n_ignored), int(ulps), cap,
                                 _p(oa, ctypes.c_int32), _p(ob, ctypes.c_int32),
                                 _p(od, ctypes.c_float))
    k = min(int(k), cap)
    return oa[:k].copy(), ob[:k].copy(), od[:k].copy()


def distances(xyz, pairs, box, orthogonal):
    """compute_distances restatement: float32 [F, P]."""
    x = _f32(xyz)
    F, N = x.shape[0], x.shape[1]
    pr = _i32(pairs).reshape(-1, 2)
    b = _box9(box)
    out = np.empty((F, pr.shape[0]), dtype=np.float32)
    lib().orc_distances(_p(x, ctypes.c_float), F, N, _p(pr, ctypes.c_int32), pr.shape[0],
                        _p(b, ctypes.c_float), int(bool(orthogonal)), _p(out, ctypes.c_float))
    return out


def min_dist(xyz, query, haystack, box, orthogonal):
    x = _f32(xyz)
    F, N = x.shape[0], x.shape[1]
    q, h = _i32(query), _i32(haystack)
    b = _box9(box)
    omin = np.empty(F, dtype=np.float32)
    oarg = np.empty(F, dtype=np.int64)
    lib().orc_min_dist(_p(x, ctypes.c_float), F, N, _p(q, ctypes.c_int32), q.size,
                       _p(h, ctypes.c_int32), h.size, _p(b, ctypes.c_float),
                       int(bool(orthogonal)), _p(omin, ctypes.c_float), _p(oarg, ctypes.c_int64))
    return omin, oarg


def max_threads():
    return int(lib().orc_max_threads())


def set_threads(n):
    """Team size of the oracle's OpenMP loops from now on (this process)."""
    lib().orc_set_threads(int(n))
