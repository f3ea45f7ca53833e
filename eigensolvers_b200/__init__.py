"""eigensolvers_b200 — B200-native (sm_100a) back-end for the NumpyVector hot path of
chem-rano/eigensolvers: `CudaVector` (AbstractVector plug-in), the device operator, and host
drivers mirroring `inexactLanczosDiagonalization` / `feastDiagonalization`.

Importing the package does not touch the GPU; the first CudaVector does.
"""
from .vector_api import AbstractVector, LINDEP_DEFAULT_VALUE  # noqa: F401

__all__ = ["AbstractVector", "LINDEP_DEFAULT_VALUE", "CudaVector", "DeviceOperator", "KroneckerSumOperator", "Runtime"]


def __getattr__(name):
    if name == "CudaVector":
        from .cudaVector import CudaVector
        return CudaVector
    if name == "DeviceOperator":
        from .operator import DeviceOperator
        return DeviceOperator
    if name == "KroneckerSumOperator":
        from .kronecker import KroneckerSumOperator
        return KroneckerSumOperator
    if name == "Runtime":
        from .runtime import Runtime
        return Runtime
    raise AttributeError(name)
