"""Shared fixtures.  Markers: `gpu` = needs a CUDA device (the parity tests proper, run on a B200)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (sm_100a); run with `-m gpu` on a B200")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    import gtcrn_oracle
    return gtcrn_oracle


@pytest.fixture(scope="session", params=["dns3", "rand7"])
def weights(request, oracle):
    return request.param, oracle.load_weights_npz(os.path.join(GOLDEN, f"weights_{request.param}.npz"))


@pytest.fixture(scope="session")
def weights_dns3(oracle):
    return oracle.load_weights_npz(os.path.join(GOLDEN, "weights_dns3.npz"))


@pytest.fixture(scope="session")
def ref_outputs():
    return dict(np.load(os.path.join(GOLDEN, "ref_outputs.npz")))


@pytest.fixture(scope="session")
def pairs():
    return dict(np.load(os.path.join(GOLDEN, "pairs.npz")))


@pytest.fixture(scope="session")
def metrics_golden():
    """waveform pairs with the reference's own SI-SNR / SDR values (tests/golden/make_golden.py make_metrics)"""
    return dict(np.load(os.path.join(GOLDEN, "metrics.npz")))


@pytest.fixture(scope="session")
def b200_lib():
    """Build (if stale) and load the C-ABI library; the GPU tests go through it."""
    from gtcrn_micro_b200 import build, lib
    build.build()
    return lib.load()
