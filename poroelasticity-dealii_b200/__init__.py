"""B200-native fixed-stress-split poroelastic time step (drop-in for the hot path of
ishovkun/poroelasticity-dealii).  The numerical product is ``lib/libporoel.so`` (CUDA, sm_100a) behind
``include/poroel.h``; this package only holds the ctypes bindings and the host-side mirror of the
reference's solver interface."""
from . import capi, fss, inputs  # noqa: F401
