import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(__file__)):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def default_p():
    """Default case parameters as Model._check_inputs leaves them (with K_b_fn, G, K_b)."""
    import numpy as np

    from crt1d_b200 import cases

    p = cases.load_default_case(60)
    G_fn = p["G_fn"]
    p["K_b_fn"] = lambda psi_: G_fn(psi_) / np.cos(psi_)
    p["G"] = G_fn(p["psi"])
    p["K_b"] = p["K_b_fn"](p["psi"])
    return p


@pytest.fixture(autouse=True)
def _restore_kernel_tuning():
    """Set up before `monkeypatch`, so finalised after its undo: make the library re-read the CRT1D_B200_*
    variables once a test's overrides are gone (they are cached at load time; see util.tune)."""
    yield
    from crt1d_b200 import _lib

    if _lib._lib is not None:
        _lib._lib.crt1d_reload_tuning()
