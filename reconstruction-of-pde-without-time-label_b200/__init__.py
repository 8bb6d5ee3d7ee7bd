"""blindno_b200 -- B200-native (sm_100a) NIO-FNO hot path behind the reference's module surface.

Layout:
  csrc/          hand-written CUDA kernels + the C ABI (include/blindno_b200.h)
  lib/           libblindno_b200.so (built in-tree by build.py; not committed)
  _lib.py        ctypes binding of the C ABI
  ops.py         autograd ops over the C ABI (no CPU fallback)
  surface/       FNO1d/FNO2d/SpectralConv*/NIOFP*_FNO with the reference's names and state_dict layout
  dropin/        per-directory FNOModules.py / NIOModules.py / ... shims for the unchanged scripts
  parallel.py    data-parallel step: flat gradient buffer + one NCCL all-reduce + fused Adam
"""
from . import _lib  # noqa: F401
from . import ops  # noqa: F401

__version__ = "0.1.0"
