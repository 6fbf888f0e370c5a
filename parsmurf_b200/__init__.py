"""parsmurf_b200 -- B200-native (sm_100a) hyperSMURF partition pipeline.

Layers:  C-ABI CUDA library (csrc/, include/parsmurf_b200.h)  <-  C++ host mirror of the reference's
hyperSMURF / Partition / SamplerKNN / rfRanger / Curves classes (host/)  <-  this thin ctypes binding.
No CPU fallback: compute entry points raise PsmError(PSM_E_CUDA) without a CUDA device.
"""
from ._capi import Context, PsmError, build, lib, KERNEL_CLASSES, LIB_PATH  # noqa: F401

__all__ = ["Context", "PsmError", "build", "lib", "KERNEL_CLASSES", "LIB_PATH"]
