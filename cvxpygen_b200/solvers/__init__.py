"""Solver-plugin classes (boundary b1): the attribute sets the reference's writer reads from a SolverInterface."""
from .admm_cuda import ADMMCUDAInterface, Setting      # noqa: F401
from .ipm_cuda import IPMCUDAInterface                # noqa: F401
