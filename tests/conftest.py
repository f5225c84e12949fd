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


@pytest.fixture(scope="session")
def weights():
    from nunet_b200.weights import load_default_weights
    try:
        return load_default_weights()
    except FileNotFoundError as e:
        pytest.skip(str(e))


@pytest.fixture(scope="session")
def golden_io():
    return dict(np.load(os.path.join(GOLDEN, "oracle_io_lstm.npz")))


@pytest.fixture(scope="session")
def tflite_weights():
    """Weights of the reference's shipped nutls_lstm.tflite (int8 tensors dequantised), role-named."""
    from nunet_b200.weights import load_tflite_lstm_weights
    try:
        return load_tflite_lstm_weights()
    except FileNotFoundError as e:
        pytest.skip(str(e))


@pytest.fixture(scope="session")
def golden_o2():
    return dict(np.load(os.path.join(GOLDEN, "o2_lstm.npz")))


@pytest.fixture(scope="session")
def ddb_weights():
    """Weights of the reference's shipped nutls.tflite (dilated-dense baseline), dequantised, role-named."""
    from nunet_b200.weights import load_ddb_weights
    try:
        return load_ddb_weights()
    except FileNotFoundError as e:
        pytest.skip(str(e))


@pytest.fixture(scope="session")
def golden_o2_ddb():
    return dict(np.load(os.path.join(GOLDEN, "o2_ddb.npz")))
