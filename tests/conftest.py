"""pytest plumbing: the `gpu` marker, import paths for the oracle (checker) and the C-ABI binding (product)."""
import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "f9-juce-resampler-studio_b200")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _load(name, path):
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (test infrastructure)."""
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def f9():
    """ctypes binding of libf9dsp.so (the product)."""
    return _load("f9dsp", os.path.join(PKG, "py", "f9dsp.py"))


@pytest.fixture(scope="session")
def ctx(f9):
    """One f9_context on cuda:0.  Creation fails loudly without a GPU: there is no CPU fallback."""
    c = f9.Context(0)
    yield c
    c.close()
