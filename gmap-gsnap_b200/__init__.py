"""B200-native banded gap-fill DP for GMAP/GSNAP (libdynprog_cuda) -- host-side mirror."""
from .api import *  # noqa: F401,F403
