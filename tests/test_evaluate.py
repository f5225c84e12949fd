"""Evaluation caller (test_interface.py:51-110 + dataloader.mapping_testset): host logic on CPU, the full run on GPU."""
import os

import numpy as np
import pytest

from nunet_b200 import evaluate as ev


def _make_set(tmp_path, n=3200 + 77):
    rng = np.random.default_rng(0)
    t = np.arange(n) / 16000.0
    for name, f0 in (("aaaa0001", 200.0), ("bbbb0002", 310.0)):
        clean = 0.5 * np.sin(2 * np.pi * f0 * t)
        ev.write_wav(str(tmp_path / f"{name}.wav"), clean, 16000)
        for snr in (0, 10):
            noise = rng.standard_normal(n) * 0.5 / np.sqrt(2) * 10 ** (-snr / 20)
            ev.write_wav(str(tmp_path / f"{name}_{snr}.wav"), np.clip(clean + noise, -1, 1), 16000)
    return str(tmp_path)


def test_wav_roundtrip_and_mapping(tmp_path):
    d = _make_set(tmp_path)
    x, fs = ev.read_wav(os.path.join(d, "aaaa0001.wav"))
    assert fs == 16000 and abs(np.abs(x).max() - 0.5) < 1e-3
    mt = ev.MappingTestset(d, d)
    assert mt.noisy_files == ["aaaa0001_0.wav", "aaaa0001_10.wav", "bbbb0002_0.wav", "bbbb0002_10.wav"]
    assert mt.get_snr_index() == ["0", "10", "0", "10"]
    clean, noisy = mt.mapping_data()
    assert all(n.shape[1] % 256 == 0 and n.shape[0] == 1 for n in noisy)          # hop alignment pad (dataloader.py:150-152)
    assert all(len(c) == n.shape[1] for c, n in zip(clean, noisy))
    assert all(np.abs(n).max() <= 1.0 and n.dtype == np.float32 for n in noisy)
    assert noisy[0][0, :256 - (3200 + 77) % 256].max() == noisy[0][0, 0]          # the pad sits in FRONT


def test_report_buckets_with_identity_model(tmp_path):
    d = _make_set(tmp_path)
    rep = ev.evaluate(lambda x, training=False: x, d, d, out_dir=str(tmp_path / "out"))
    assert set(rep) == {"0", "10", "all"} and rep["0"]["n"] == 2 and rep["all"]["n"] == 4
    assert abs(rep["0"]["input"] - rep["0"]["output"]) < 1e-6                      # identity model
    assert rep["10"]["input"] > rep["0"]["input"] + 5.0
    assert os.path.exists(str(tmp_path / "out" / "aaaa0001_0.wav"))


@pytest.mark.gpu
def test_evaluate_on_reference_excerpt_gpu(weights, tmp_path):
    """The reference's own noisy / clean pair (committed 1.5 s excerpt) through models.NUTLS_LSTM: SI-SDR must rise."""
    from nunet_b200 import models
    from nunet_b200.options import default_options
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "wav_excerpt.npz"))
    ev.write_wav(str(tmp_path / "40hc020i.wav"), g["clean"].astype(np.float64) / 32768.0, 16000)
    ev.write_wav(str(tmp_path / "40hc020i_0.wav"), g["noisy"].astype(np.float64) / 32768.0, 16000)
    model = models.NUTLS_LSTM(default_options()).build_model().load_weights(weights)
    rep = ev.evaluate(model, str(tmp_path), str(tmp_path))
    assert rep["0"]["n"] == 1 and rep["0"]["output"] > rep["0"]["input"] + 8.0, rep
