"""ORACLE-ONLY stand-in for the third-party ``mdtraj`` package (test infrastructure).

MDTraj (unpinned dependency of the reference, /root/reference/setup.cfg:27) is not
installable in this environment.  This package provides just enough of its API for
the UNMODIFIED reference Python (``/root/reference/contact_map``) to be imported
and run in the build container, so that the reference's own combinatorics
validate the oracle restatement and generate the golden vectors under
``tests/golden``.  It is never imported by the product package and never
travels onto the product path.

The two numerical entry points, ``compute_neighborlist`` and
``compute_distances``, call the C restatement in ``oracle/cm_oracle.c``.
"""
from . import element                                   # noqa: F401
from .topology import Topology, Atom, Residue, Chain    # noqa: F401
from .trajectory import Trajectory, load, load_pdb, join  # noqa: F401
from .geometry import compute_neighborlist, compute_distances  # noqa: F401


__version__ = "0.0-oracle-shim"
from . import core                                      # noqa: F401,E402
