"""Pins of the oracle against the reference's DEPLOYED computation: the shipped `.tflite` flatbuffers executed op by
op (oracle/tflite_graph.py).  TensorFlow / TFLite cannot be installed here, so this is the strongest available
statement of "what the reference computes" (SURVEY 4 item 6, 8c)."""
import os

import numpy as np
import pytest
import torch

REF_TFLITE = "/root/reference/dnn_model/tflite"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF_TFLITE), reason="reference checkout not mounted (GPU box)")


def test_source_restatement_equals_shipped_graph(tflite_weights, golden_o2):
    """O1 (offline restatement, zero padding, frame_div32 CTFA) on the dequantised weights == the flatbuffer run
    frame by frame with carried history, to fp32 round-off.  Validates padding, taps, both sub-pixel shuffles,
    down/up-sampling orientation, LayerNorm, PReLU, LSTM gate order, skip wiring and zero-history == zero-padding."""
    from oracle.nunet_oracle import Oracle
    o = Oracle(tflite_weights, ctfa_mode="frame_div32")
    mag = torch.from_numpy(golden_o2["mag"])
    with torch.no_grad():
        est = o.net(mag[None, :, :, None]).squeeze().numpy()
    ref = golden_o2["model_out"]
    assert ref.max() > 20.0
    assert np.abs(est - ref).max() <= 1e-4, np.abs(est - ref).max()


def test_frame_graph_restatement_carries_the_same_history(tflite_weights, golden_o2):
    """The one-frame restatement (converter_proposed.py:188-867) driven like the interpreter ends with the same
    history tensors as the flatbuffer."""
    from oracle.nunet_oracle import Oracle
    o = Oracle(tflite_weights, ctfa_mode="frame_div32")
    mag = golden_o2["mag"]
    state = o.zero_state(1)
    with torch.no_grad():
        for t in range(mag.shape[0]):
            feed = {"input": torch.from_numpy(mag[t].reshape(1, 1, 256, 1))}
            feed.update({k.replace("_cur", "_prev"): v for k, v in state.items()})
            res = o.frame_step(feed)
            out = res.pop("model_out")
            state = res
    assert np.abs(out.reshape(256).numpy() - golden_o2["model_out"][-1]).max() <= 1e-4
    for k in ("msfe6_ee_cur1", "msfe4_dd2_cur3", "state_h", "msfe5_en_c"):
        ref = golden_o2[f"state_{k}"]
        assert np.abs(state[k].numpy().reshape(ref.shape) - ref).max() <= 1e-4, k


@needs_ref
def test_reader_reproduces_the_graph_inventory():
    """Operator histograms and signature sizes of SURVEY 2.1 / Appendix A.2."""
    import collections
    from nunet_b200.tflite_reader import read_tflite
    g = read_tflite(f"{REF_TFLITE}/nutls_lstm.tflite")
    h = collections.Counter(o.op for o in g.operators)
    assert len(g.operators) == 3066 and h["CONV_2D"] == 172 and h["TRANSPOSE_CONV"] == 6 and h["FULLY_CONNECTED"] == 39
    assert h["PRELU"] == 117 and h["LOGISTIC"] == 63 and h["AVERAGE_POOL_2D"] == 12
    assert g.signatures[0].key == "nutls_lstm_sm" and len(g.signatures[0].inputs) == len(g.signatures[0].outputs) == 131
    assert sum(1 for t in g.tensors if t.data is not None and t.dtype == np.int8) == 211
    g2 = read_tflite(f"{REF_TFLITE}/nutls.tflite")
    h2 = collections.Counter(o.op for o in g2.operators)
    assert len(g2.operators) == 4353 and h2["CONV_2D"] == 354 and h2["SPACE_TO_BATCH_ND"] == 65
    assert g2.signatures[0].key == "nutls" and len(g2.signatures[0].inputs) == 209


@needs_ref
def test_h5_and_tflite_weights_agree_within_half_a_quantisation_step(weights, tflite_weights):
    """|w_h5 - q*scale| <= scale/2 for every tensor: pins the role naming incl. the Dense remap (proposed.py:47-63)."""
    from nunet_b200.tflite_reader import read_tflite
    assert set(weights) == set(tflite_weights)
    worst = 0.0
    for k, w in weights.items():
        d = np.abs(w - tflite_weights[k]).max()
        worst = max(worst, d / max(np.abs(w).max(), 1e-12))
        assert d <= 0.0041 * np.abs(w).max() + 1e-7, (k, d, np.abs(w).max())   # 1/254 of the tensor range + fp slack
    assert worst > 1e-4    # the int8 tensors really are quantised


@needs_ref
def test_graph_executor_matches_committed_fixture(golden_o2):
    """The fixture is reproducible from the reference file (first 6 frames; the executor is slow)."""
    from oracle.tflite_graph import TFLiteGraph, stream_frames
    g = TFLiteGraph(f"{REF_TFLITE}/nutls_lstm.tflite")
    est = stream_frames(g, golden_o2["mag"][:6])
    assert np.abs(est - golden_o2["model_out"][:6]).max() <= 1e-5


# ---------------------------------------------------------------------------------------------------
# dilated-dense baseline (models/nunet_tls.py, shipped nutls.tflite)
def test_ddb_restatement_equals_shipped_graph(ddb_weights, golden_o2_ddb):
    """Group / concat channel order, time taps t-d & t, frequency taps f-d, f, f+d, history = last d rows, and the
    shared down-sampling weight of the shipped file: 72 frames cover the d = 32 layer."""
    from oracle.nunet_oracle import Oracle
    o = Oracle(ddb_weights, ctfa_mode="frame_div32", variant="ddb")
    mag = torch.from_numpy(golden_o2_ddb["mag"])
    with torch.no_grad():
        est = o.net(mag[None, :, :, None]).squeeze().numpy()
    ref = golden_o2_ddb["model_out"]
    assert np.abs(ref).max() > 1.0
    assert np.abs(est - ref).max() <= 1e-4, np.abs(est - ref).max()


def test_ddb_streaming_restatement_equals_offline(ddb_weights, golden_o2_ddb):
    """One-frame form with the reference's history tensors (`*_ddb_prevK`, converter_nunet_tls.py:373-411)."""
    from oracle.nunet_oracle import Oracle
    o = Oracle(ddb_weights, ctfa_mode="frame_div32", variant="ddb")
    mag = torch.from_numpy(golden_o2_ddb["mag"][:40])
    with torch.no_grad():
        ref = o.net(mag[None, :, :, None]).squeeze()
        st, outs = None, []
        for t in range(mag.shape[0]):
            nxt = {}
            if st is None:      # zero history of the right shapes: run once with st_out only
                o.net(mag[None, t:t + 1, :, None], st_out=nxt)
                st = {k.replace("_cur", "_prev"): torch.zeros_like(v) for k, v in nxt.items()}
                nxt = {}
            outs.append(o.net(mag[None, t:t + 1, :, None], st_in=st, st_out=nxt).reshape(256))
            st = {k.replace("_cur", "_prev"): v for k, v in nxt.items()}
    assert (torch.stack(outs) - ref).abs().max() <= 1e-4


def test_ddb_weight_set_structure(ddb_weights):
    from nunet_b200.weights import expected_ddb_shapes, validate
    validate(ddb_weights, expected_ddb_shapes())
    assert sum(v.size for v in ddb_weights.values()) == 2830462
    k = ddb_weights["msfe4_down_sampling/kernel"]
    assert (ddb_weights["msfe4_down_sampling2/kernel"] == k).all() and (ddb_weights["msfe4_down_sampling3/kernel"] == k).all()


def test_hybrid_quantisers_follow_the_reference_kernels():
    """Asymmetric/SymmetricQuantizeFloats (portable_tensor_utils.cc) on hand-checked rows."""
    from oracle.tflite_graph import TFLiteGraph
    x = torch.tensor([[-1.0, 0.0, 3.0], [0.0, 0.0, 0.0], [2.0, 4.0, 8.0]])
    q, s = TFLiteGraph._quantize_rows(x, asymmetric=True)
    # row 0: scale 4/255, zero point round(-128 + 1/scale) = -64 -> q = (-128, -64, 127); returned minus the zero point
    assert np.allclose(s.numpy(), [4 / 255, 1.0, 8 / 255]) and q[0].tolist() == [-64.0, 0.0, 191.0] and q[1].tolist() == [0.0, 0.0, 0.0]
    # all-positive row: zero point -128; 4 lands on -128 + 127.5 = -0.5, which std::round takes away from zero to -1
    assert q[2].tolist() == [64.0, 127.0, 255.0]
    q, s = TFLiteGraph._quantize_rows(x, asymmetric=False)
    assert np.allclose(s.numpy(), [3 / 127, 1.0, 8 / 127]) and q[0].tolist() == [-42.0, 0.0, 127.0] and q[2].tolist() == [32.0, 64.0, 127.0]


@needs_ref
def test_hybrid_emulation_of_the_shipped_graph(golden_o2):
    """what the TFLite runtime does with the dynamic-range file (int8 activations per call in 170 CONV_2D and
    35 FULLY_CONNECTED ops) stays close to the float graph -- and nowhere near the 1e-3 bar the engine is held to: on the
    reference's noisy wav the deployed arithmetic is ~42 dB below the float output (max |d| ~0.7 on a peak of 48)."""
    from oracle.tflite_graph import TFLiteGraph, stream_frames
    mag, ref = golden_o2["mag"][:24], golden_o2["model_out"][:24]
    g = TFLiteGraph(os.path.join(REF_TFLITE, "nutls_lstm.tflite"), hybrid=True)
    n_hybrid = sum(1 for op in g.g.operators if op.op in ("CONV_2D", "FULLY_CONNECTED") and g._is_int8(op.inputs[1]))
    assert n_hybrid == 205
    out = stream_frames(g, mag)
    snr = 10 * np.log10((ref ** 2).sum() / ((out - ref) ** 2).sum())
    assert 25.0 < snr < 60.0, snr
    assert np.abs(out - ref).max() > 1e-2
    assert np.array_equal(out, stream_frames(TFLiteGraph(os.path.join(REF_TFLITE, "nutls_lstm.tflite"), hybrid=True), mag))
