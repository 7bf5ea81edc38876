import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    return binding.oracle()


@pytest.fixture(scope="session")
def reference():
    """The reference's compiled C path; tests that pin the oracle against it skip when it is absent."""
    from oracle import binding
    ref = binding.reference()
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference in this environment)")
    return ref


@pytest.fixture
def experiments():
    """Routes lib.call() through libhevcasm_b200_exp.so for the duration of one test: the same sources built with
    -DHEVCASM_EXPERIMENTS, i.e. with the HEVCASM_* switches that pin one code path (the product library has none) and the
    measured-but-not-adopted kernel variants."""
    from hevcasm_b200 import lib
    lib.use_experiments(True)
    yield
    lib.use_experiments(False)
