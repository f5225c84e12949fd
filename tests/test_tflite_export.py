"""Artifact emission, deployment side (converter_proposed.py:877-912): the exporter turns a float weight set
into the reference's one-frame `.tflite`.  Pinned against the reference's own artefact: quantising the reference's float
checkpoint (`log/saved_model/nutls_lstm.h5`, here as the committed blob) must reproduce the reference's SHIPPED
`tflite/nutls_lstm.tflite` byte for byte -- its SHA-256 is the golden value below (recorded by make_golden.py)."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_artifacts.json")))


def test_export_of_the_float_checkpoint_is_the_shipped_tflite_byte_for_byte(weights, tmp_path):
    from nunet_b200.tflite_export import export_lstm_tflite
    p = str(tmp_path / "out.tflite")
    info = export_lstm_tflite(weights, p)
    assert info == {"int8_tensors": 211, "float_tensors": 561, "bytes": GOLD["nutls_lstm.tflite"]["bytes"]}
    assert hashlib.sha256(open(p, "rb").read()).hexdigest() == GOLD["nutls_lstm.tflite"]["sha256"]
    ref = "/root/reference/dnn_model/tflite/nutls_lstm.tflite"
    if os.path.exists(ref):                  # build container: compare with the file itself
        assert open(p, "rb").read() == open(ref, "rb").read()


def test_exported_file_round_trips_and_runs_in_the_flatbuffer_executor(weights, tmp_path):
    """Other weights than the shipped ones: a perturbed set is written, read back within half a quantisation step, and the
    written GRAPH, executed op by op (oracle/tflite_graph.py), equals the source restatement on the read-back weights."""
    from nunet_b200.models import NUTLS_LSTM
    from nunet_b200.options import default_options
    from nunet_b200.tflite_reader import read_tflite
    from nunet_b200.weights import lstm_weights_from_tflite
    from oracle.nunet_oracle import Oracle
    from oracle.tflite_graph import TFLiteGraph
    rng = np.random.default_rng(5)
    w = {k: (v * (1.0 + 0.05 * rng.standard_normal(v.shape))).astype(np.float32) for k, v in weights.items()}
    p = str(tmp_path / "perturbed.tflite")
    fm = NUTLS_LSTM(default_options()).tflite_model().load_weights(w)
    info = fm.convert_to_tflite(p)                     # the converter surface; host-only
    assert info["int8_tensors"] == 211
    g = read_tflite(p)
    assert g.signatures[0].key == "nutls_lstm_sm" and len(g.signatures[0].inputs) == 131 and len(g.operators) == 3066
    back = lstm_weights_from_tflite(p)
    for k, v in w.items():
        if k.endswith("upsampling/bias"):
            continue
        step = np.abs(v).max() / 127.0
        assert np.abs(back[k] - v).max() <= 0.5 * step + 1e-7, k
    o2 = TFLiteGraph(p)
    o1 = Oracle(back, ctfa_mode="frame_div32")
    feed = {k: torch.zeros(s) for k, s in o2.input_shapes().items()}
    st = o1.zero_state(1)
    mags = np.abs(rng.standard_normal((3, 256))).astype(np.float32) * 4
    for t in range(3):
        feed["input"] = torch.from_numpy(mags[t].reshape(1, 1, 256, 1))
        out2 = o2.run(feed)
        f1 = {"input": feed["input"]}
        f1.update({k.replace("_cur", "_prev"): v for k, v in st.items()})
        with torch.no_grad():
            out1 = o1.frame_step(f1)
        assert float((out2["model_out"] - out1["model_out"]).abs().max()) <= 1e-4
        st = {k: v for k, v in out1.items() if k != "model_out"}
        for k in list(feed):
            if k != "input":
                feed[k] = out2[k.replace("_prev", "_cur")]


@pytest.mark.gpu
def test_engine_serves_the_exported_file(weights, tmp_path):
    """Interpreter(model_path=<exported .tflite>) through the CUDA engine == the flatbuffer executor on the same file."""
    from nunet_b200.interpreter import Interpreter
    from nunet_b200.tflite_export import export_lstm_tflite
    from oracle.tflite_graph import TFLiteGraph
    p = str(tmp_path / "out.tflite")
    export_lstm_tflite(weights, p)
    it = Interpreter(model_path=p)
    it.allocate_tensors()
    run = it.get_signature_runner("nutls_lstm_sm")
    o2 = TFLiteGraph(p)
    feed = {k: np.zeros(s, np.float32) for k, s in o2.input_shapes().items()}
    rng = np.random.default_rng(9)
    for t in range(3):
        feed["input"] = (np.abs(rng.standard_normal((1, 1, 256, 1))) * 5).astype(np.float32)
        ours = run(**feed)
        ref = o2.run({k: torch.from_numpy(np.asarray(v)) for k, v in feed.items()})
        assert np.abs(ours["model_out"] - ref["model_out"].numpy()).max() <= 1e-3
        for k in list(feed):
            if k != "input":
                feed[k] = ours[k.replace("_prev", "_cur")]
