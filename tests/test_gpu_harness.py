"""GPU: random differential test through the function-select boundary (oracle/ref_harness.c, test infrastructure): the self-test's
per-block cases on fresh random inputs through the reference's compiled C slots (HEVCASM_C_REF | HEVCASM_C_OPT) and through this
library's HEVCASM_CUDA slots; every output digest must agree.  The analogue of the reference's own loop over instruction sets
(hevcasm_test.c:110-137) for an instruction set that lives on another device."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libhevcasm_cref.so")
GPU_SO = os.path.join(ROOT, "hevcasm_b200", "libhevcasm_b200.so")


@pytest.mark.parametrize("first_seed", [1, 1000])
def test_reference_slots_vs_cuda_slots(first_seed):
    if not (os.path.exists(HARNESS) and os.path.exists(REF_SO)):
        pytest.skip("oracle/_ref was never built (no /root/reference where this tree was prepared)")
    r = subprocess.run([HARNESS, REF_SO, GPU_SO, "24", str(first_seed)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "0 differences" in r.stdout
