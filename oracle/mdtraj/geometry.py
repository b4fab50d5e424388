"""``compute_neighborlist`` / ``compute_distances`` of the oracle's mdtraj stand-in.

Python-wrapper behaviour restated from MDTraj's published API: periodic iff the
trajectory carries a unit cell; ``compute_distances`` decides orthogonality with
``np.allclose(unitcell_angles, 90)``; out-of-range pair indices raise ValueError.
"""
import os
import sys

import numpy as np

try:
    from oracle import clib
except ImportError:   # imported as top-level ``mdtraj`` with only oracle/ on sys.path
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from oracle import clib


def compute_neighborlist(traj, cutoff, frame=0, periodic=True):
    if traj.topology is None:
        raise ValueError("traj must have a topology defined")
    box = None
    if periodic and traj._have_unitcell:
        box = np.asarray(traj.unitcell_vectors[frame], dtype=np.float32)
    xyz = np.asarray(traj.xyz[frame], dtype=np.float32)
    return clib.neighborlist(xyz, box, cutoff)


def compute_distances(traj, atom_pairs, periodic=True, opt=True):
    xyz = np.asarray(traj.xyz, dtype=np.float32)
    pairs = np.asarray(list(atom_pairs), dtype=np.int32).reshape(-1, 2)
    if not np.all(np.logical_and(pairs < traj.n_atoms, pairs >= 0)):
        raise ValueError("atom_pairs must be between 0 and %d" % traj.n_atoms)
    if len(pairs) == 0:
        return np.zeros((len(xyz), 0), dtype=np.float32)
    if periodic and traj._have_unitcell:
        box = np.asarray(traj.unitcell_vectors, dtype=np.float32)
        orthogonal = np.allclose(traj.unitcell_angles, 90)
        return clib.distances(xyz, pairs, box, orthogonal)
    return clib.distances(xyz, pairs, None, True)
