"""Shared helpers for the parity tests (the oracle is the checker, never the product)."""
import numpy as np


def system_arrays(system, query=None, haystack=None):
    """Per-atom arrays (all atoms used): residue, chain, role bits."""
    top = system.topology
    n = top.n_atoms
    atom_res, atom_chain = top.atom_residue_chain()
    role = np.zeros(n, dtype=np.uint8)
    q = np.arange(n) if query is None else np.asarray(query)
    h = np.arange(n) if haystack is None else np.asarray(haystack)
    role[q] |= 1
    role[h] |= 2
    return atom_res.astype(np.int32), atom_chain.astype(np.int32), role


def system_arrays_of_topology(top):
    """Per-atom arrays of a bare topology, every atom query and haystack."""
    atom_res, atom_chain = top.atom_residue_chain()
    return atom_res.astype(np.int32), atom_chain.astype(np.int32), np.full(top.n_atoms, 3, dtype=np.uint8)


def dense_from_export(idx, cnt, n):
    out = np.zeros(n * n, dtype=np.uint32)
    out[idx] = cnt
    return out.reshape(n, n)


def keys_to_frames(keys):
    keys = np.asarray(keys, dtype=np.uint64)
    frame = (keys >> np.uint64(40)).astype(np.int64)
    lo = ((keys >> np.uint64(20)) & np.uint64(0xFFFFF)).astype(np.int64)
    hi = (keys & np.uint64(0xFFFFF)).astype(np.int64)
    return frame, lo, hi
