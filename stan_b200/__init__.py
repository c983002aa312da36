"""stan_b200 -- B200-native GLM log-density + gradient backend for Stan's NUTS hot path.

Only what the path needs lives here: csrc/ (sm_100a kernels + the C ABI of include/b200glm.h),
cpp/ (the C++ shim that plugs into the reference's model / integrator interfaces) and the
Python mirror of that interface used by the tests and bench.py.
"""
from .model import GLMModel, DomainError, InvalidArgument, CudaError  # noqa: F401
from .synth import make_glm_data, theta_points  # noqa: F401

__all__ = ["GLMModel", "DomainError", "InvalidArgument", "CudaError", "make_glm_data", "theta_points"]
