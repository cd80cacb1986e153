from .admm_cuda import ADMMCUDAInterface, Setting   # noqa: F401
