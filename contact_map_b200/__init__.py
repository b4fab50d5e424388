"""contact_map_b200 -- B200-native (sm_100a) drop-in for the hot path of dwhswenson/contact_map.

Public names follow the reference's export list (/root/reference/contact_map/__init__.py:8-27)
for the classes on the hot path.  Importing the package never touches the GPU; the first
computation loads ``libcmb200.so`` (built in-tree with nvcc) and creates a CUDA context.  There
is no CPU fallback.
"""
from .contact_count import ContactCount                                    # noqa: F401
from .contact_map import (                                                 # noqa: F401
    ContactMap, ContactFrequency, ContactDifference, ContactObject,
    AtomMismatchedContactDifference, ResidueMismatchedContactDifference,
    OverrideTopologyContactDifference, residue_neighborhood,
)
from .contact_trajectory import (                                          # noqa: F401
    ContactTrajectory, MutableContactTrajectory, RollingContactFrequency, WindowedIterator,
)
from .min_dist import NearestAtoms, MinimumDistanceCounter                 # noqa: F401
from .runner import (                                                      # noqa: F401
    ChunkedContactFrequency, DaskContactFrequency, DaskContactTrajectory,
)
from .concurrence import (                                                 # noqa: F401
    Concurrence, AtomContactConcurrence, ResidueContactConcurrence,
)
from . import frequency_task                                               # noqa: F401
from .trajectory import Trajectory, Topology, load, load_pdb               # noqa: F401

__version__ = "0.1.0"
