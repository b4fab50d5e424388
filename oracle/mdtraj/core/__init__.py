"""``mdtraj.core`` namespace of the oracle's stand-in (element, topology, trajectory)."""
from .. import element                      # noqa: F401
from .. import topology                     # noqa: F401
from .. import trajectory                   # noqa: F401
