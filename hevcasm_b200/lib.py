"""Loader for the product's native library (CUDA kernels + C ABI).  There is no fallback: if libhevcasm_b200.so is
missing or does not export the whole ABI this raises - nothing in this package computes on the CPU."""
import ctypes as C
import os

from .abi import BATCH_ABI, GPU_ONLY_ABI, HOST_ABI, P

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhevcasm_b200.so")
# the same sources built with -DHEVCASM_EXPERIMENTS: measured-but-not-adopted kernel variants + HEVCASM_* environment switches.
# Only tools/ and the parity tests that pin one code path load it (use_experiments()); bench.py and smoke() never do.
EXP_LIB_PATH = os.path.join(_HERE, "libhevcasm_b200_exp.so")

_lib = None
_libs = {}


class HevcasmError(RuntimeError):
    pass


def use_experiments(on=True):
    """Route lib.call() through libhevcasm_b200_exp.so (on) or back through the product library (off)."""
    global _lib
    _lib = _bind(EXP_LIB_PATH if on else LIB_PATH)
    return _lib


def load():
    global _lib
    if _lib is None:
        _lib = _bind(LIB_PATH)
    return _lib


def _bind(path):
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise HevcasmError(
            f"{path} is missing: build it with `make -C hevcasm_b200/csrc` (or __graft_entry__.build()). "
            "hevcasm_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    for table in (BATCH_ABI, GPU_ONLY_ABI):
        for name, args in table.items():
            fn = getattr(lib, "hevcasm_" + name)  # AttributeError if the symbol is missing: fail loudly
            fn.argtypes = list(args) + [P]        # + void *stream
            fn.restype = C.c_int
    for name, args in HOST_ABI.items():
        fn = getattr(lib, "hevcasm_" + name)
        fn.argtypes = list(args)
        fn.restype = C.c_int
    lib.hevcasm_cuda_context_create.argtypes = [C.c_int, C.c_size_t]
    lib.hevcasm_cuda_context_create.restype = P
    lib.hevcasm_cuda_context_destroy.argtypes = [P]
    lib.hevcasm_cuda_context_destroy.restype = None
    lib.hevcasm_cuda_context_stream.argtypes = [P]
    lib.hevcasm_cuda_context_stream.restype = P
    lib.hevcasm_cuda_host_alloc.argtypes = [C.c_size_t]
    lib.hevcasm_cuda_host_alloc.restype = P
    lib.hevcasm_cuda_host_alloc_near.argtypes = [C.c_size_t, C.c_int]
    lib.hevcasm_cuda_host_alloc_near.restype = P
    lib.hevcasm_cuda_device_numa_node.argtypes = [C.c_int]
    lib.hevcasm_cuda_device_numa_node.restype = C.c_int
    lib.hevcasm_cuda_host_free.argtypes = [P]
    lib.hevcasm_cuda_host_free.restype = None
    lib.hevcasm_cuda_error_string.argtypes = [C.c_int]
    lib.hevcasm_cuda_error_string.restype = C.c_char_p
    lib.hevcasm_cuda_launch_count.argtypes = []
    lib.hevcasm_cuda_launch_count.restype = C.c_ulonglong
    lib.hevcasm_instruction_set_support.argtypes = []
    lib.hevcasm_instruction_set_support.restype = C.c_int
    _libs[path] = lib
    return lib


def check(code):
    if code != 0:
        raise HevcasmError(f"hevcasm_b200: error {code}: {load().hevcasm_cuda_error_string(code).decode()}")


def call(name, *args, stream=None):
    """Enqueue hevcasm_<name>(*args, stream); raises HevcasmError on a non-zero return."""
    fn = getattr(load(), "hevcasm_" + name)
    check(fn(*args, stream))


def call_host(name, ctx, *args):
    """hevcasm_<name>(ctx, *args) for the host-memory forms (synchronous); raises HevcasmError on a non-zero return."""
    fn = getattr(load(), "hevcasm_" + name)
    check(fn(ctx, *args))


class Context:
    """hevcasm_cuda_context: streams + device arena for the *_host entry points."""

    def __init__(self, device=0, arena_bytes=1 << 30):
        self.handle = load().hevcasm_cuda_context_create(device, arena_bytes)
        if not self.handle:
            raise HevcasmError(f"hevcasm_cuda_context_create(device={device}, arena_bytes={arena_bytes}) failed")

    def close(self):
        if self.handle:
            load().hevcasm_cuda_context_destroy(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def pinned_array(shape, dtype, device=None):
    """numpy array over page-locked host memory from hevcasm_cuda_host_alloc (device=None) or, bound to the NUMA node of a GPU,
    hevcasm_cuda_host_alloc_near; the allocation is released when the array (and every view of it) has been garbage-collected."""
    import weakref

    import numpy as np
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    lib = load()
    p = lib.hevcasm_cuda_host_alloc(max(n, 1)) if device is None else lib.hevcasm_cuda_host_alloc_near(max(n, 1), int(device))
    if not p:
        raise HevcasmError(f"hevcasm_cuda_host_alloc({n}) failed")
    buf = (C.c_uint8 * max(n, 1)).from_address(p)
    weakref.finalize(buf, lib.hevcasm_cuda_host_free, p)   # views keep `buf` alive through their .base chain
    return np.frombuffer(buf, dtype=np.uint8, count=n).view(dtype).reshape(shape)


def launch_count():
    return int(load().hevcasm_cuda_launch_count())
