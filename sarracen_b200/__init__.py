"""sarracen_b200 -- B200-native (sm_100a) SPH particle -> grid scatter.

A from-scratch implementation of the one data-parallel hot path of Sarracen
(`sarracen.interpolate`): hand-written CUDA behind a C ABI
(include/sphb200.h, sarracen_b200/csrc), plugged in behind Sarracen's own
backend interface (`B200Backend`).  See DESIGN.md and INTEGRATION.md.
"""
from .backend import B200Backend, install, resolve_weight_function  # noqa: F401
from .frame import DeviceFrame, ParticleFrame  # noqa: F401
from .kernels import (BaseKernel, CubicSplineKernel, QuarticSplineKernel,  # noqa: F401
                      QuinticSplineKernel, WeightFunction)

__version__ = "0.1.0"
