"""Golden vectors: SHA-256 digests of the reference's own C path outputs on seeded planes (tests/golden/golden.json, produced
by tests/golden/make_golden.py where /root/reference exists).  The CPU oracle must reproduce them (CPU suite), and so must
the CUDA library through its C ABI (-m gpu) - neither test needs the reference at run time."""
import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import cases  # noqa: E402

with open(os.path.join(HERE, "golden", "golden.json")) as f:
    GOLDEN = json.load(f)["sha256"]
CASES = dict(cases.cases())


def test_golden_file_covers_every_case():
    assert set(GOLDEN) == set(CASES) and len(GOLDEN) >= 39


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_reference_digest(oracle, name):
    assert CASES[name](cases.CpuBackend(oracle)) == GOLDEN[name]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_reproduces_reference_digest(name):
    assert CASES[name](cases.GpuBackend()) == GOLDEN[name]
