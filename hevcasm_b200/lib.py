"""Loader for the product's native library (CUDA kernels + C ABI).  There is no fallback: if libhevcasm_b200.so is
missing or does not export the whole ABI this raises - nothing in this package computes on the CPU."""
import ctypes as C
import os

from .abi import BATCH_ABI, GPU_ONLY_ABI, P

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhevcasm_b200.so")

_lib = None


class HevcasmError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HevcasmError(
            f"{LIB_PATH} is missing: build it with `make -C hevcasm_b200/csrc` (or __graft_entry__.build()). "
            "hevcasm_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for table in (BATCH_ABI, GPU_ONLY_ABI):
        for name, args in table.items():
            fn = getattr(lib, "hevcasm_" + name)  # AttributeError if the symbol is missing: fail loudly
            fn.argtypes = list(args) + [P]        # + void *stream
            fn.restype = C.c_int
    lib.hevcasm_cuda_error_string.argtypes = [C.c_int]
    lib.hevcasm_cuda_error_string.restype = C.c_char_p
    lib.hevcasm_cuda_launch_count.argtypes = []
    lib.hevcasm_cuda_launch_count.restype = C.c_ulonglong
    lib.hevcasm_instruction_set_support.argtypes = []
    lib.hevcasm_instruction_set_support.restype = C.c_int
    _lib = lib
    return lib


def check(code):
    if code != 0:
        raise HevcasmError(f"hevcasm_b200: error {code}: {load().hevcasm_cuda_error_string(code).decode()}")


def call(name, *args, stream=None):
    """Enqueue hevcasm_<name>(*args, stream); raises HevcasmError on a non-zero return."""
    fn = getattr(load(), "hevcasm_" + name)
    check(fn(*args, stream))


def launch_count():
    return int(load().hevcasm_cuda_launch_count())
