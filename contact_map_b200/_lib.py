"""ctypes binding of ``libcmb200.so`` (the C ABI declared in ``include/cmb200.h``).

The library is built in-tree (``contact_map_b200/libcmb200.so``) by ``build()``;
there is no CPU fallback: if the library is missing and cannot be built, or no
CUDA device is present when a context is created, a RuntimeError is raised.
"""
import ctypes
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
# CMB200_SO selects an alternative build (kernel-tuning experiments); the default is in-tree
SO_PATH = os.environ.get("CMB200_SO") or os.path.join(_HERE, "libcmb200.so")
_SOURCES = [os.path.join(_HERE, "csrc", f) for f in (
    "cmb_api.cu", "cmb_math.cuh", "cmb_cells.cuh", "cmb_frames_small.cuh",
    "cmb_frames_large.cuh", "cmb_aux_kernels.cuh", "cmb_verlet.cuh")] + [os.path.join(_ROOT, "include", "cmb200.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

_LIB = None

EXPORTS = [
    "cmb_version", "cmb_last_error", "cmb_ctx_create", "cmb_ctx_destroy", "cmb_set_system",
    "cmb_n_residues", "cmb_residue_map", "cmb_table_sizes", "cmb_path_kind", "cmb_accumulate",
    "cmb_accumulate_host", "cmb_accumulate_host_gather", "cmb_frame_contacts", "cmb_sort_keys", "cmb_export_nonzero",
    "cmb_boundary_pairs", "cmb_set_option", "cmb_get_stat", "cmb_gather_atoms", "cmb_min_dist",
    "cmb_nearest", "cmb_synth_frames", "cmb_set_table_mode", "cmb_hash_stats", "cmb_export_hashed",
    "cmb_hash_rehash", "cmb_hash_overflow_buffer", "cmb_hash_add", "cmb_export_sorted", "cmb_group_contacts",
    "cmb_min_dist_pairs", "cmb_frame_contacts_host", "cmb_keys_to_csr", "cmb_screen_audit", "cmb_device_alloc", "cmb_device_free", "cmb_device_zero",
    "cmb_copy_to_host", "cmb_copy_to_device", "cmb_nccl_unique_id", "cmb_nccl_comm_create", "cmb_nccl_comm_destroy",
    "cmb_reduce", "cmb_unpack_sorted", "cmb_unpack_pairs", "cmb_wait_for_stream",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def needs_build():
    if os.environ.get("CMB200_SO"):
        return False
    if not os.path.exists(SO_PATH):
        return True
    so_time = os.path.getmtime(SO_PATH)
    return any(os.path.getmtime(src) > so_time for src in _SOURCES if os.path.exists(src))


def build(force=False, verbose=False):
    """Compile the CUDA extension for sm_100a with nvcc (cross-compiles without a GPU)."""
    if not force and not needs_build():
        return SO_PATH
    nvcc = _nvcc()
    if nvcc is None:
        raise RuntimeError("libcmb200.so is missing or stale and nvcc was not found; "
                           "contact_map_b200 has no CPU fallback")
    tmp = SO_PATH + ".tmp%d" % os.getpid()
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp, _SOURCES[0]]
    env = dict(os.environ)
    env.pop("CC", None)        # the image exports a wrapper gcc; let nvcc pick the system one
    env.pop("CXX", None)
    proc = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), proc.stderr[-4000:]))
    os.replace(tmp, SO_PATH)
    if verbose:
        print(proc.stderr)
    return SO_PATH


def lib():
    """Load (building first if necessary) the shared library and declare signatures."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if needs_build():
        build()
    L = ctypes.CDLL(SO_PATH)
    vp, i32, i64, u64, dbl, flt = (ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint64,
                                   ctypes.c_double, ctypes.c_float)
    L.cmb_version.restype = i32
    L.cmb_last_error.restype = ctypes.c_char_p
    L.cmb_last_error.argtypes = [vp]
    L.cmb_ctx_create.restype = i32
    L.cmb_ctx_create.argtypes = [i32, ctypes.POINTER(vp)]
    L.cmb_ctx_destroy.restype = None
    L.cmb_ctx_destroy.argtypes = [vp]
    L.cmb_set_system.restype = i32
    L.cmb_set_system.argtypes = [vp, i32, vp, vp, vp, dbl, i32]
    L.cmb_n_residues.restype = i32
    L.cmb_n_residues.argtypes = [vp]
    L.cmb_residue_map.restype = i32
    L.cmb_residue_map.argtypes = [vp, vp]
    L.cmb_table_sizes.restype = i32
    L.cmb_table_sizes.argtypes = [vp, ctypes.POINTER(i64), ctypes.POINTER(i64)]
    L.cmb_path_kind.restype = i32
    L.cmb_path_kind.argtypes = [vp]
    L.cmb_accumulate.restype = i32
    L.cmb_accumulate.argtypes = [vp, vp, vp, i64, vp, vp, vp]
    L.cmb_accumulate_host.restype = i32
    L.cmb_accumulate_host.argtypes = [vp, vp, vp, i64, i64, vp, vp]
    L.cmb_accumulate_host_gather.restype = i32
    L.cmb_accumulate_host_gather.argtypes = [vp, vp, i32, vp, vp, i64, i64, i32, vp, vp]
    L.cmb_set_table_mode.restype = i32
    L.cmb_set_table_mode.argtypes = [vp, i32, i32]
    L.cmb_hash_stats.restype = i32
    L.cmb_hash_stats.argtypes = [vp, ctypes.POINTER(i64), ctypes.POINTER(i64), ctypes.POINTER(i64)]
    L.cmb_export_hashed.restype = i32
    L.cmb_export_hashed.argtypes = [vp, vp, i32, i64, vp, vp, vp, vp]
    L.cmb_hash_overflow_buffer.restype = i32
    L.cmb_hash_overflow_buffer.argtypes = [vp, vp, i64]
    L.cmb_hash_add.restype = i32
    L.cmb_hash_add.argtypes = [vp, vp, i32, vp, vp, i64, i32, vp]
    L.cmb_export_sorted.restype = i32
    L.cmb_export_sorted.argtypes = [vp, vp, i64, i32, i64, vp, vp, vp]
    L.cmb_unpack_sorted.restype = i32
    L.cmb_unpack_sorted.argtypes = [vp, vp, vp, i64, i64, vp, vp, vp]
    L.cmb_wait_for_stream.restype = i32
    L.cmb_wait_for_stream.argtypes = [vp, vp]
    L.cmb_unpack_pairs.restype = i32
    L.cmb_unpack_pairs.argtypes = [vp, vp, vp, i64, i64, i64, i64, vp, vp, vp, vp, vp]
    L.cmb_group_contacts.restype = i32
    L.cmb_group_contacts.argtypes = [vp, vp, vp, i64, i32, vp, i32, vp, vp, i32, i32, flt, vp, vp]
    L.cmb_hash_rehash.restype = i32
    L.cmb_hash_rehash.argtypes = [vp, vp, i32, vp, i32, vp]
    L.cmb_frame_contacts.restype = i32
    L.cmb_frame_contacts.argtypes = [vp, vp, vp, i64, i64, vp, i64, vp, i64, vp, vp, vp, vp]
    L.cmb_frame_contacts_host.restype = i32
    L.cmb_frame_contacts_host.argtypes = [vp, vp, i32, vp, vp, i64, i64, i64, i32, vp, i64, vp, i64, vp]
    L.cmb_keys_to_csr.restype = i32
    L.cmb_keys_to_csr.argtypes = [vp, vp, i64, i64, i64, vp, vp, vp, vp, vp]
    L.cmb_screen_audit.restype = i32
    L.cmb_screen_audit.argtypes = [vp, vp, vp, i64, vp, vp]
    L.cmb_device_alloc.restype = i32
    L.cmb_device_alloc.argtypes = [vp, u64, ctypes.POINTER(vp)]
    L.cmb_device_free.restype = i32
    L.cmb_device_free.argtypes = [vp, vp]
    L.cmb_device_zero.restype = i32
    L.cmb_device_zero.argtypes = [vp, vp, u64, vp]
    L.cmb_copy_to_host.restype = i32
    L.cmb_copy_to_host.argtypes = [vp, vp, vp, u64, vp]
    L.cmb_copy_to_device.restype = i32
    L.cmb_copy_to_device.argtypes = [vp, vp, vp, u64, vp]
    L.cmb_nccl_unique_id.restype = i32
    L.cmb_nccl_unique_id.argtypes = [vp, vp]
    L.cmb_nccl_comm_create.restype = i32
    L.cmb_nccl_comm_create.argtypes = [vp, i32, vp, i32, ctypes.POINTER(vp)]
    L.cmb_nccl_comm_destroy.restype = i32
    L.cmb_nccl_comm_destroy.argtypes = [vp, vp]
    L.cmb_reduce.restype = i32
    L.cmb_reduce.argtypes = [vp, vp, vp, i64, i32, vp]
    L.cmb_sort_keys.restype = i32
    L.cmb_sort_keys.argtypes = [vp, vp, i64, vp]
    L.cmb_export_nonzero.restype = i32
    L.cmb_export_nonzero.argtypes = [vp, vp, i64, i64, vp, vp, vp, vp]
    L.cmb_boundary_pairs.restype = i32
    L.cmb_boundary_pairs.argtypes = [vp, vp, vp, i64, i64, i32, vp, i64, vp, vp]
    L.cmb_set_option.restype = i32
    L.cmb_set_option.argtypes = [vp, ctypes.c_char_p, i64]
    L.cmb_get_stat.restype = i32
    L.cmb_get_stat.argtypes = [vp, ctypes.c_char_p, ctypes.POINTER(i64)]
    L.cmb_gather_atoms.restype = i32
    L.cmb_gather_atoms.argtypes = [vp, vp, i64, i32, vp, i32, vp, vp]
    L.cmb_min_dist.restype = i32
    L.cmb_min_dist.argtypes = [vp, vp, vp, i64, i32, vp, i32, vp, i32, i32, vp, vp, vp]
    L.cmb_min_dist_pairs.restype = i32
    L.cmb_min_dist_pairs.argtypes = [vp, vp, vp, i64, i32, vp, i64, i32, vp, vp, vp]
    L.cmb_nearest.restype = i32
    L.cmb_nearest.argtypes = [vp, vp, vp, i32, dbl, i32, i32, vp, vp, vp, vp, vp, vp]
    L.cmb_synth_frames.restype = i32
    L.cmb_synth_frames.argtypes = [vp, vp, vp, i32, vp, u64, flt, i32, i64, i64, vp, vp]
    _LIB = L
    return L
