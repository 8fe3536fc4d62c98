"""pytest configuration: `gpu` marker, repo on sys.path, fixtures for the CPU test double and golden data."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def have_reference():
    return os.path.isdir(os.path.join(REFERENCE, "klongpy"))


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        meta = json.load(f)
    meta["compiled_arrays"] = dict(np.load(os.path.join(GOLDEN, "golden_compiled.npz"), allow_pickle=False))
    meta["verb_arrays"] = dict(np.load(os.path.join(GOLDEN, "golden_verbs.npz"), allow_pickle=False))
    return meta


def to_ir(x):
    """JSON lists -> the tuple IR of klongpy/compiler.py."""
    if isinstance(x, list):
        return tuple(to_ir(v) for v in x)
    return x


@pytest.fixture(scope="session")
def built_lib():
    """Build (if needed) and load libkgb200.so; no GPU required to load it."""
    from klongpy_b200.csrc import build
    build.build()
    from klongpy_b200 import _lib
    return _lib.load()


@pytest.fixture(scope="session")
def cuda_backend():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from klongpy_b200 import B200BackendProvider
    return B200BackendProvider(device="cuda:0")


@pytest.fixture(scope="session")
def fake_backend():
    """Provider on the CPU test double (host-logic tests only)."""
    import fake_lib
    dev = fake_lib.install()
    from klongpy_b200 import B200BackendProvider
    return B200BackendProvider(device=dev)
