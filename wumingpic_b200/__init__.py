"""wumingpic_b200 -- B200-native backend for WumingPIC's per-timestep PIC loop.

The product is the C-ABI shared library ``lib/libwuming_b200.so`` (hand-written sm_100a CUDA,
``include/wuming_b200.h``).  This package is the host-side mirror of the reference's Fortran
module interface for that path (``particle__solv``, ``field__fdtd_i``, ``bc__particle_x``,
``bc__particle_yz``, ``sort__bucket`` ...) over ctypes.  There is no CPU fallback: importing the
backend without the built library, or creating a context without a CUDA device, fails loudly.
"""
from .backend import Backend, ShockParams, WmError, load_library, para_range, weibel_constants  # noqa: F401
from .shock import inject_counts  # noqa: F401
from .mpi_set import SlabLayout  # noqa: F401

__all__ = ["Backend", "ShockParams", "inject_counts", "WmError", "load_library", "para_range", "weibel_constants", "SlabLayout"]
