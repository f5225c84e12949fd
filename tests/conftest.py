import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run by the driver with -m gpu)")


def _have_cuda() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a CPU host skips the gpu-marked tests instead of failing them with 'no CUDA device'.
    An explicit `-m gpu` run on a box without a GPU still fails loudly (nothing may pass silently there)."""
    if _have_cuda() or "gpu" in (config.getoption("-m") or "").replace("not gpu", ""):
        return
    skip = pytest.mark.skip(reason="no CUDA device on this host")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def _load_or_fail(loader, what):
    """Weight blobs are committed under nunet_b200/data/.  A missing blob must never make the parity suite vanish:
    on a GPU box it is a hard failure; on a CPU-only host (oracle tests) it is a skip."""
    try:
        return loader()
    except FileNotFoundError as e:
        if _have_cuda():
            pytest.fail(f"{what} missing on a GPU box: {e}")
        pytest.skip(str(e))


@pytest.fixture(scope="session")
def weights():
    from nunet_b200.weights import load_default_weights
    return _load_or_fail(load_default_weights, "trained LSTM weight blob")


@pytest.fixture(scope="session")
def golden_io():
    return dict(np.load(os.path.join(GOLDEN, "oracle_io_lstm.npz")))


@pytest.fixture(scope="session")
def tflite_weights():
    """Weights of the reference's shipped nutls_lstm.tflite (int8 tensors dequantised), role-named."""
    from nunet_b200.weights import load_tflite_lstm_weights
    return _load_or_fail(load_tflite_lstm_weights, "nutls_lstm.tflite weight blob")


@pytest.fixture(scope="session")
def golden_o2():
    return dict(np.load(os.path.join(GOLDEN, "o2_lstm.npz")))


@pytest.fixture(scope="session")
def ddb_weights():
    """Weights of the reference's shipped nutls.tflite (dilated-dense baseline), dequantised, role-named."""
    from nunet_b200.weights import load_ddb_weights
    return _load_or_fail(load_ddb_weights, "nutls.tflite (DDB) weight blob")


@pytest.fixture(scope="session")
def golden_o2_ddb():
    return dict(np.load(os.path.join(GOLDEN, "o2_ddb.npz")))
