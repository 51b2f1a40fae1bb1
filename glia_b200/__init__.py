"""glia_b200 -- B200-native (sm_100a) reaction-diffusion forward/adjoint hot path of GLIA.

Everything numerical lives in ``lib/libglia_rd.so`` (hand-written CUDA behind the C ABI of
``include/glia_rd.h``).  This package is the ctypes binding plus the host-side mirror of
the reference's operator classes; it has no CPU fallback.
"""
from ._capi import GliaRdError, load_library  # noqa: F401
from .rd import RDHandle  # noqa: F401
