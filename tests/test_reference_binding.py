"""INTEGRATION.md B made executable: `contact_map_b200.reference_binding.bind(package)` subclasses
the package's ContactFrequency / ContactTrajectory through the two hooks the reference's own Dask
runner overrides, using nothing but the C ABI (ctypes over include/cmb200.h, no torch).

* GPU: bound to THIS package's classes (same hook contract; the reference cannot travel to the GPU
  box) -- results against the golden vectors the unmodified reference generated.
* GPU + reference importable (a maintainer's machine): bound to the real reference.
* CPU (this container, where /root/reference is importable through the oracle's mdtraj stand-in):
  the bound classes ARE subclasses of the reference's, override exactly the two hooks, and fail
  loudly -- no CPU fallback -- when there is no GPU.
"""
import collections
import os
import sys
import warnings

import numpy as np
import pytest

import contact_map_b200 as cm
from contact_map_b200 import reference_binding
from tests import goldens

warnings.simplefilter("ignore", PendingDeprecationWarning)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _import_reference():
    """The reference package, if this machine has it (MDTraj, or the oracle's stand-in for it)."""
    ref_root = "/root/reference"
    if not os.path.isdir(os.path.join(ref_root, "contact_map")):
        pytest.skip("the reference package is not on this machine")
    for path in (os.path.join(ROOT, "oracle"), ref_root):
        if path not in sys.path:
            sys.path.insert(0, path)
    cwd = os.getcwd()
    os.chdir(ref_root)                 # (its version discovery looks for setup.cfg relative to the cwd)
    try:
        return pytest.importorskip("contact_map")
    finally:
        os.chdir(cwd)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["pdb_c045_n2", "syn_tric_sub", "syn_ortho_n0", "syn_open_all"])
def test_bound_classes_reproduce_the_reference_goldens(engine, name):
    B200ContactFrequency, B200ContactTrajectory = reference_binding.bind(cm)
    case = goldens.load(name)
    traj = goldens.trajectory(case)
    kw = goldens.call_kwargs(case)
    m = B200ContactFrequency(traj, **kw)
    assert isinstance(m, cm.ContactFrequency)
    assert m._atom_contacts == goldens.counter(case["atom_pairs"], case["atom_counts"])
    assert m._residue_contacts == goldens.counter(case["res_pairs"], case["res_counts"])
    assert m.atom_contacts.counter == goldens.counter(case["atom_freq_pairs"], case["atom_freq_values"])
    ct = B200ContactTrajectory(traj, **kw)
    want_atoms = goldens.frame_sets(case["frame_atom_ptr"], case["frame_atom_pairs"])
    want_res = goldens.frame_sets(case["frame_res_ptr"], case["frame_res_pairs"])
    assert len(ct) == len(traj)
    for f, frame in enumerate(ct):
        assert frame._atom_contacts == collections.Counter(want_atoms[f])
        assert frame._residue_contacts == collections.Counter(want_res[f])
    # the bound class and the package's own class agree object for object
    assert m == cm.ContactFrequency(traj, **kw)


@pytest.mark.gpu
def test_bound_to_the_real_reference(engine):
    ref = _import_reference()
    import mdtraj as md
    B200ContactFrequency, B200ContactTrajectory = reference_binding.bind(ref)
    traj = md.load(os.path.join("/root/reference", "contact_map", "tests", "trajectory.pdb"))
    for kw in ({"cutoff": 0.075, "n_neighbors_ignored": 0}, {"cutoff": 0.45, "n_neighbors_ignored": 2}):
        want = ref.ContactFrequency(traj, **kw)
        got = B200ContactFrequency(traj, **kw)
        assert got._atom_contacts == want._atom_contacts and got._residue_contacts == want._residue_contacts
        assert got == want
        assert B200ContactTrajectory(traj, **kw) == ref.ContactTrajectory(traj, **kw)


def test_binding_subclasses_the_reference_and_never_falls_back():
    ref = _import_reference()
    import torch
    B200ContactFrequency, B200ContactTrajectory = reference_binding.bind(ref)
    assert issubclass(B200ContactFrequency, ref.ContactFrequency)
    assert issubclass(B200ContactTrajectory, ref.ContactTrajectory)
    assert B200ContactFrequency._build_contact_map is not ref.ContactFrequency._build_contact_map
    assert B200ContactTrajectory._build_contacts is not ref.ContactTrajectory._build_contacts
    # nothing else of the reference is touched
    overridden = {k for k in vars(B200ContactFrequency) if not k.startswith(("__", "_abc"))} | \
                 {k for k in vars(B200ContactTrajectory) if not k.startswith(("__", "_abc"))}
    assert overridden == {"_build_contact_map", "_build_contacts"}
    # the module talks to the C ABI only
    src = open(reference_binding.__file__).read()
    assert "import torch" not in src and "engine" not in src.replace("contact_map_b200/engine.py", "")
    if torch.cuda.is_available():
        return
    import mdtraj as md
    traj = md.load(os.path.join("/root/reference", "contact_map", "tests", "trajectory.pdb"))
    with pytest.raises(reference_binding.CmbError):
        B200ContactFrequency(traj, cutoff=0.075, n_neighbors_ignored=0)
    with pytest.raises(reference_binding.CmbError):
        B200ContactTrajectory(traj, cutoff=0.075, n_neighbors_ignored=0)
