"""particles_b200 -- the per-step SMC hot path of nchopin/particles on a B200.

Hand-written sm_100a CUDA kernels (csrc/, C-ABI in include/smcb.h) behind the
reference's own plugin surface: ``SMC``, ``FeynmanKac``, ``state_space_models``,
``distributions``, ``resampling``, ``collectors``.  See DESIGN.md / INTEGRATION.md.
"""
from .core import SMC, FeynmanKac  # noqa: F401
from .device import seed  # noqa: F401

__version__ = "0.1.0"
