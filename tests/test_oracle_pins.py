"""CPU pins of the oracle (oracle/nunet_oracle.py) against everything the reference offers for this path:
literal window tables, history-tensor tables, weight-file structure, the noisy/clean wav pair, offline ==
streaming self-consistency, and the committed golden outputs.  """
import json
import os

import numpy as np
import pytest
import torch

from oracle.nunet_oracle import (Oracle, hann_window_periodic, inverse_stft_window, min_max_norm, si_sdr,
                                 state_shapes)

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_windows_match_java_tables():
    """RTSE_NUTLS_LSTM.java:62-63 hold the TF hann window (ends forced to 1e-7, interpreter_proposed.py:22) and
    the inverse-STFT window as float32 literals."""
    j = json.load(open(os.path.join(GOLDEN, "java_windows.json")))
    w = hann_window_periodic(512).numpy().copy()
    w[0] = w[-1] = 1e-7
    assert np.abs(w - np.array(j["window"], np.float32)).max() < 2e-7
    inv = inverse_stft_window(512, 256).numpy()
    assert np.abs(inv - np.array(j["inverse_window"], np.float32)).max() < 1e-6


def test_state_tables_match_reference():
    ref = json.load(open(os.path.join(GOLDEN, "state_shapes_lstm.json")))
    ref.pop("input"), ref.pop("model_out")
    ours = state_shapes()
    assert set(ours) == set(ref) and len(ours) == 130
    for k, v in ref.items():
        assert tuple(v) == ours[k], k
    assert sum(int(np.prod(v)) for v in ours.values()) == 205090          # 0.820 MB per stream
    specs = json.load(open(os.path.join(GOLDEN, "input_specs_lstm.json")))
    assert len(specs) == 131 and specs[0][0] == "input"
    for name, shape in specs[1:]:
        cur = name.replace("_prev", "_cur")
        assert tuple(shape[1:]) == ours[cur][1:], name


def test_product_state_table_agrees():
    from nunet_b200.state_table import STATE_FLOATS, STATE_SHAPES
    assert STATE_SHAPES == state_shapes() and STATE_FLOATS == 205090


def test_weight_set_structure(weights):
    from nunet_b200.weights import expected_lstm_shapes, pack_blob, unpack_blob, validate
    validate(weights, expected_lstm_shapes())
    assert sum(v.size for v in weights.values()) == 2832910               # parameter count of the .h5
    w2, variant = unpack_blob(pack_blob(weights))
    assert variant == 0 and all(np.array_equal(w2[k], weights[k]) for k in weights)


@pytest.mark.skipif(not os.path.exists("/root/reference/dnn_model/log/saved_model/nutls_lstm.h5"),
                    reason="reference checkout not mounted")
def test_h5_reader_and_dense_remap():
    from nunet_b200.h5_reader import read_h5
    from nunet_b200.weights import lstm_weights_from_h5
    raw = read_h5("/root/reference/dnn_model/log/saved_model/nutls_lstm.h5")
    assert len(raw) == 772 and len({k.split("/")[1] for k in raw}) == 180
    w = lstm_weights_from_h5("/root/reference/dnn_model/log/saved_model/nutls_lstm.h5")
    # proposed.py:47-63 mis-names six decoder Dense layers; Keras loads topologically
    assert np.array_equal(w["msfe3_de_dense/kernel"], raw["/msfe6_de_dense/msfe6_de_dense/kernel:0"])
    assert w["msfe6_de_dense/kernel"].shape == (21, 128)


def test_oracle_matches_committed_golden(weights, golden_io):
    with torch.no_grad():
        for mode in ("causal_avg32", "frame_div32"):
            y, est = Oracle(weights, ctfa_mode=mode).forward_wav(golden_io["wav"][:1])
            assert np.abs(est.numpy() - golden_io[f"est_{mode}"][:1]).max() < 2e-4
            assert np.abs(y.numpy() - golden_io[f"wav_{mode}"][:1]).max() < 2e-4
    # the two CTFA modes are different functions
    assert np.abs(golden_io["est_causal_avg32"] - golden_io["est_frame_div32"]).max() > 0.05


def test_enhances_reference_excerpt(weights):
    g = np.load(os.path.join(GOLDEN, "wav_excerpt.npz"))
    noisy, clean = g["noisy"] / 32768.0, g["clean"] / 32768.0
    x = min_max_norm(noisy).astype(np.float32)
    with torch.no_grad():
        y, est = Oracle(weights).forward_wav(x[None])
    y = y[0].numpy()
    before, after = si_sdr(clean[:len(y)], x[:len(y)]), si_sdr(clean[:len(y)], y)
    assert after > before + 8.0, (before, after)


def test_streaming_equals_offline(weights):
    """Zero history == zero time padding; with carried CTFA history the one-frame graph reproduces the offline
    graph, and in frame_div32 mode the offline restatement equals the frame graph exactly."""
    from nunet_b200.synth import synth_clips
    T = 36
    wav = synth_clips(1, 512 + 256 * (T - 1), first_clip=3)
    for mode, hist in (("frame_div32", None), ("causal_avg32", {})):
        o = Oracle(weights, ctfa_mode=mode)
        mags, _ = o.stft(torch.from_numpy(wav))
        mag = mags[:, :, 1:]
        with torch.no_grad():
            ref = o.net(mag[..., None]).squeeze(-1)
            state = o.zero_state(1)
            outs = []
            for t in range(T):
                feed = {"input": mag[:, t].reshape(1, 1, 256, 1)}
                feed.update({k.replace("_cur", "_prev"): v for k, v in state.items()})
                res = o.frame_step_hist(feed, hist) if hist is not None else o.frame_step(feed)
                outs.append(res.pop("model_out").reshape(1, 256))
                state = res
        out = torch.stack(outs, dim=1)
        assert float((out - ref).abs().max()) < 2e-4, mode


def test_fp64_twin_bounds_fp32_error(weights):
    from nunet_b200.synth import synth_clips
    wav = synth_clips(1, 512 + 256 * 15, first_clip=9)
    with torch.no_grad():
        _, e32 = Oracle(weights).forward_wav(wav)
        _, e64 = Oracle(weights, dtype=torch.float64).forward_wav(wav)
    assert float((e32.double() - e64).abs().max()) < 2e-4
