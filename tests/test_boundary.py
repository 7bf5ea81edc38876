"""CPU-side checks of the drop-in boundary (no GPU needed, no compute calls):
the C-ABI library loads and exports every symbol include/*.h declares; the public headers compile as C99 and their table
structs have the reference's layout; populate() with no implemented instruction set leaves every slot empty (there is no
CPU fallback) and with HEVCASM_CUDA fills exactly the slots the reference's C path serves."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")
REF_INC = "/root/reference/src/lib"
HEVCASM_CUDA = 1 << 9
ALL_CPU_BITS = (1 << 9) - 1

TABLES = {  # populate function -> number of function-pointer slots in the table struct
    "sad": 12, "sad_multiref": 16 * 16 + 1, "ssd": 5, "pred_uni_8to8": 2 * 9 * 2 * 2, "pred_bi_8to8": 2 * 5 * 2, "transform": 5,
    "inverse_transform_add": 5, "quantize": 1, "quantize_inverse": 1, "quantize_reconstruct": 4, "hadamard_satd": 3}


@pytest.fixture(scope="module")
def lib():
    from hevcasm_b200 import lib as L
    return L.load()


def declared_symbols():
    names = set()
    for fn in os.listdir(INC):
        text = open(os.path.join(INC, fn)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        text = re.sub(r"typedef[^;{]*;", "", text)   # function-type typedefs are not symbols
        names |= set(re.findall(r"HEVCASM_API\s*\*?\s*(hevcasm_\w+)\s*\(", text))
        names |= set(re.findall(r"hevcasm_test_function\s+(hevcasm_\w+)\s*;", text))
    return names


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) >= 53, sorted(names)
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing


def _populate(lib, name, mask):
    n = TABLES[name]
    table = (C.c_void_p * n)(*([0xDEAD] * n))
    fn = getattr(lib, "hevcasm_populate_" + name)
    if name == "inverse_transform_add":
        fn.argtypes = [C.c_void_p, C.c_int, C.c_int]
        fn(table, mask, 0)
    else:
        fn.argtypes = [C.c_void_p, C.c_int]
        fn(table, mask)
    return [v or 0 for v in table]


@pytest.mark.parametrize("name", sorted(TABLES))
def test_no_cpu_fallback_in_tables(lib, name):
    for mask in (0, ALL_CPU_BITS, 1, 2, 1 << 8):
        slots = _populate(lib, name, mask)
        if name == "sad_multiref":
            slots = slots[:-1] + [0]  # sadGeneric_4 is never written by the reference either; we zero it
        assert not any(slots), (name, mask)


@pytest.mark.parametrize("name", sorted(TABLES))
def test_cuda_bit_fills_the_reference_slots(lib, name):
    slots = _populate(lib, name, HEVCASM_CUDA | ALL_CPU_BITS)
    if name == "sad_multiref":
        lookup = [slots[i * 16:(i + 1) * 16] for i in range(16)]
        assert all(all(row) for row in lookup) and slots[-1] == 0
    elif name == "pred_uni_8to8":     # [taps/4-1][ceil(w/taps)][x][y]: index 0 of the width bucket is unreachable (w >= 1)
        for t in range(2):
            for b in range(9):
                cell = slots[(t * 9 + b) * 4:(t * 9 + b) * 4 + 4]
                assert all(cell) == (b >= 1) and any(cell) == (b >= 1), (t, b)
    elif name == "pred_bi_8to8":
        for t in range(2):
            for b in range(5):
                cell = slots[(t * 5 + b) * 2:(t * 5 + b) * 2 + 2]
                assert all(cell) == (b >= 1), (t, b)
    else:
        assert all(slots), name


C_PROBE = r"""
#include <stdio.h>
#include <stddef.h>
#include "hevcasm.h"
#include "sad.h"
#include "ssd.h"
#include "pred_inter.h"
#include "residual_decode.h"
#include "quantize.h"
#include "hadamard.h"
#include "diff.h"
int main(void) {
    hevcasm_table_sad t; hevcasm_table_pred_uni_8to8 u; hevcasm_table_pred_bi_8to8 b; hevcasm_table_sad_multiref m;
    printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(hevcasm_table_sad), sizeof(hevcasm_table_sad_multiref), sizeof(hevcasm_table_ssd),
           sizeof(hevcasm_table_pred_uni_8to8), sizeof(hevcasm_table_pred_bi_8to8), sizeof(hevcasm_table_transform),
           sizeof(hevcasm_table_inverse_transform_add), sizeof(hevcasm_table_quantize), sizeof(hevcasm_table_quantize_inverse),
           sizeof(hevcasm_table_quantize_reconstruct));
    printf("%td %td %td %td %td\n", (char *)hevcasm_get_sad(&t, 64, 64) - (char *)&t, (char *)hevcasm_get_sad(&t, 8, 4) - (char *)&t,
           (char *)hevcasm_get_sad(&t, 12, 16) - (char *)&t, (char *)hevcasm_get_pred_uni_8to8(&u, 4, 9, 8, 1, 0) - (char *)&u,
           (char *)hevcasm_get_pred_bi_8to8(&b, 8, 33, 8, 0, 0, 0, 2) - (char *)&b);
    printf("%td %d %d\n", (char *)hevcasm_get_sad_multiref(&m, 4, 24, 32) - (char *)&m, (int)HEVCASM_RECT(48, 64), (int)HEVCASM_AVX2);
    { hevcasm_table_hadamard_satd h; printf("%zu %td\n", sizeof h, (char *)hevcasm_get_hadamard_satd(&h, 3) - (char *)&h); }
    return 0;
}
"""


def _compile_and_run(tmp_path, inc_dirs, tag):
    src = tmp_path / f"probe_{tag}.c"
    exe = tmp_path / f"probe_{tag}"
    src.write_text(C_PROBE)
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-Wno-unused-variable", "-Wno-unused-function"] + [f"-I{d}" for d in inc_dirs] + [str(src), "-o", str(exe)]
    subprocess.run(cmd, check=True, capture_output=True)
    return subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout


def test_headers_are_c99_and_match_the_reference_layout(tmp_path):
    ours = _compile_and_run(tmp_path, [INC], "ours")
    assert ours.split()[0] == str(12 * 8)
    if os.path.isdir(REF_INC):
        theirs = _compile_and_run(tmp_path, [REF_INC], "ref")
        assert ours == theirs   # same struct sizes, same slot addresses for the same getter arguments, same macros


def test_ssd_linear_getter_has_no_cpu_fallback(lib):
    lib.hevcasm_get_ssd_linear.argtypes = [C.c_int, C.c_int]
    lib.hevcasm_get_ssd_linear.restype = C.c_void_p
    assert not lib.hevcasm_get_ssd_linear(64, ALL_CPU_BITS) and not lib.hevcasm_get_ssd_linear(64, 0)
    assert lib.hevcasm_get_ssd_linear(64, HEVCASM_CUDA)


def test_selftest_binary_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by test_gpu_tables.py")
    exe = os.path.join(ROOT, "hevcasm_b200", "hevcasm_selftest")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "no compute-capability 10.x CUDA device" in r.stdout


def test_instruction_set_support_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert lib.hevcasm_instruction_set_support() == 0
