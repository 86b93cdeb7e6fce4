"""scalaopt_b200 — B200-native evaluation of scalaopt's data-set objective (loss, J^T r, J^T J,
Hessian-vector products, multi-alpha line search) behind the reference's DataSet / MSEFunction /
DifferentiableObjectiveFunction interface.

``scalaopt_b200.native``  ctypes binding of the C ABI (``include/scalaopt_cuda.h``)
``scalaopt_b200.csrc``    CUDA sources of ``libscalaopt_cuda.so`` (sm_100a)
"""
from . import native  # noqa: F401

__all__ = ["native"]
