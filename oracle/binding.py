"""TEST INFRASTRUCTURE ONLY.  ctypes bindings of the two CPU libraries:

  * `oracle()`    - oracle/_build/liboracle.so : our restatement (hevc_oracle.c), symbols oracle_* / oracle_drv_*
  * `reference()` - oracle/_ref/libhevcasm_cref.so : the reference's own C path compiled in place from
                    /root/reference by oracle/build_ref.sh, symbols ref_* / ref_drv_*  (None when it was never built)

Both expose `.drv(name, *args, threads=1)` with the argument lists of hevcasm_b200.abi.BATCH_ABI (host pointers) and
the per-block functions as `.blk.<name>`.
"""
import ctypes as C
import os
import subprocess

from hevcasm_b200.abi import BATCH_ABI, P, PD, I, U32

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "_build", "liboracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libhevcasm_cref.so")

_BLOCK_ABI = {  # per-block functions, argument order of hevc_oracle.h
    "sad": ([P, PD, P, PD, U32], I),
    "sad_multiref_4": ([P, PD, P, PD, P, U32], None),
    "pred_uni": ([P, PD, P, PD, I, I, I, I, I], None),
    "pred_bi": ([P, PD, P, P, PD, I, I, I, I, I, I, I], None),
    "transform": ([P, P, PD, I, I], None),
    "inverse_transform_add": ([P, PD, P, PD, P, I, I], None),
    "quantize": ([P, P, I, I, I, I], I),
    "quantize_inverse": ([P, P, I, I, I], None),
    "pred_coefficient": ([I, I, I], I),
    "hadamard_satd": ([P, PD, P, PD, I], I),
    "ssd_linear": ([P, P, I], I),
}


class _Ns:
    pass


class CpuLib:
    def __init__(self, path, prefix):
        self.path, self.prefix = path, prefix
        self.lib = C.CDLL(path)
        self.blk = _Ns()
        for name, (args, res) in _BLOCK_ABI.items():
            fn = getattr(self.lib, prefix + name)
            fn.argtypes, fn.restype = args, res
            setattr(self.blk, name, fn)
        # ssd / quantize_reconstruct differ in their last argument(s) between the two libraries
        ssd = getattr(self.lib, prefix + "ssd")
        rec = getattr(self.lib, prefix + "quantize_reconstruct")
        ssd.restype, rec.restype = I, None
        if prefix == "oracle_":
            ssd.argtypes = [P, PD, P, PD, I, I]
            rec.argtypes = [P, PD, P, PD, P, I]
            self.blk.ssd = lambda a, sa, b, sb, log2: ssd(a, sa, b, sb, 1 << log2, 1 << log2)
            self.blk.quantize_reconstruct = lambda r, sr, p, sp, res, log2: rec(r, sr, p, sp, res, 1 << log2)
            tm = self.lib.oracle_transform_matrix
            tm.argtypes, tm.restype = [I, I, I, I], I
            self.blk.transform_matrix = tm
        else:
            ssd.argtypes = [P, PD, P, PD, I]
            rec.argtypes = [P, PD, P, PD, P, I]
            self.blk.ssd, self.blk.quantize_reconstruct = ssd, rec
            self.lib.ref_drv_set_avx2_sad.argtypes = [I]
        self._drv = {}
        for name, args in BATCH_ABI.items():
            fn = getattr(self.lib, prefix + "drv_" + name)
            fn.argtypes, fn.restype = list(args) + [I], None
            self._drv[name] = fn
        getattr(self.lib, prefix + "drv_init")()

    def drv(self, name, *args, threads=1):
        self._drv[name](*args, threads)


_cache = {}


def build(force=False):
    """(Re)build liboracle.so and, when /root/reference exists, the reference's C path."""
    if force or not os.path.exists(ORACLE_SO) or (os.path.isdir("/root/reference") and not os.path.exists(REF_SO)):
        subprocess.run(["make", "-C", _HERE, "all"], check=True, capture_output=True)


def oracle():
    if "o" not in _cache:
        build()
        _cache["o"] = CpuLib(ORACLE_SO, "oracle_")
    return _cache["o"]


def reference():
    """The reference's own compiled C path, or None if oracle/_ref was never built (no /root/reference here)."""
    if "r" not in _cache:
        build()
        _cache["r"] = CpuLib(REF_SO, "ref_") if os.path.exists(REF_SO) else None
    return _cache["r"]


def ptr(arr, offset_elems=0):
    """host address of element `offset_elems` of a numpy array"""
    return C.c_void_p(arr.ctypes.data + offset_elems * arr.itemsize)
