import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """build the CUDA library (nvcc cross-compiles without a GPU) and the oracle once per session;
    on the GPU box the prebuilt .so files travel with the snapshot and are reused"""
    import importlib
    build = importlib.import_module("legenddsp.jl_b200.build")
    try:
        build.build_library()
    except Exception:
        if not os.path.exists(build.OUT):
            raise
    from oracle import oracle as O
    try:
        O.build()
    except Exception:
        if not os.path.exists(O._SO):
            raise
    yield


@pytest.fixture(scope="session")
def L():
    import legenddsp.jl_b200 as L
    return L


@pytest.fixture(scope="session")
def O():
    from oracle import oracle as O
    return O


@pytest.fixture(scope="session")
def example_params(L, O):
    """sample-domain params of the reference's example config (test/test_dsp_icpc.jl:50-161), tau = 500 us,
    default filter parameters, built with the ORACLE's own coefficient builders"""
    return L.resolve_icpc_params(L.example_config(), L.us(500.0), builders=O.OracleBuilders())


@pytest.fixture(scope="session")
def handle(L):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    h = L.Handle(0)
    yield h
    h.close()
