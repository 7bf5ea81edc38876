"""hevcasm_b200 - B200-native (CUDA sm_100a) build of the HEVCasm inner-loop kernels.

The product is the native library `libhevcasm_b200.so` (hevcasm_b200/csrc, C ABI in include/*.h).  This package is
only the thin Python binding used by tests and bench.py: `lib.call("<entry point>", ...)` with raw device pointers.
"""
from .abi import HEVCASM_RECT  # noqa: F401
from . import lib  # noqa: F401

HEVCASM_CUDA = 1 << 9
