"""Regenerates the committed fixtures in tests/golden/ from the reference checkout (/root/reference).

Run in the build container only (the GPU box has no /root/reference):  python tests/golden/make_golden.py
Fixtures:
  java_windows.json        literal analysis / synthesis window tables of RTSE_NUTLS_LSTM.java:62-63
  state_shapes_lstm.json   history-tensor table of interpreter_proposed.py:36-198
  input_specs_lstm.json    TensorSpec names/shapes of converter_proposed.py:26-187
  wav_excerpt.npz          first 1.5 s (int16) of data/40hc020i_0.wav (noisy) and data/40hc020i.wav (clean)
  o2_lstm.npz              outputs of the reference's shipped nutls_lstm.tflite executed by oracle/tflite_graph.py
  o2_ddb.npz               the same for nutls.tflite (dilated-dense baseline)
  reference_artifacts.json SHA-256 / size of the reference's shipped .tflite / .h5 files (exporter golden values)
  oracle_io_lstm.npz       oracle outputs with the reference .h5 weights on seeded inputs (regression pin +
                           expected values for the GPU parity tests)
"""
import ast
import glob
import json
import os
import re
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def java_windows():
    path = glob.glob(f"{REF}/mobile_app/app/src/main/java/**/RTSE_NUTLS_LSTM.java", recursive=True)[0]
    src = open(path).read()
    out = {}
    for name in ("window", "inverse_window"):
        m = re.search(r"this\.%s = new double\[\]\{([^}]*)\}" % name, src)
        out[name] = [float(v) for v in m.group(1).split(",")]
        assert len(out[name]) == 512
    json.dump(out, open(f"{HERE}/java_windows.json", "w"))


def state_tables():
    src = open(f"{REF}/dnn_model/interpreter_proposed.py").read()
    shapes = {}
    for name, shp in re.findall(r"['\"](\w+)['\"]\s*:\s*np\.zeros\(\(([\d, ]+)\)", src):
        shapes[name] = [int(v) for v in shp.split(",")]
    json.dump(shapes, open(f"{HERE}/state_shapes_lstm.json", "w"), indent=0)
    src = open(f"{REF}/dnn_model/converter_proposed.py").read()
    specs = re.findall(r"tf\.TensorSpec\(shape=\[([\w, ]+)\], dtype=tf\.float32, name='(\w+)'\)", src)
    json.dump([[n, [None if v.strip() == "None" else int(v) for v in s.split(",")]] for s, n in specs],
              open(f"{HERE}/input_specs_lstm.json", "w"), indent=0)


def wav_excerpt():
    from oracle.wavio import read_wav
    noisy, fs = read_wav(f"{REF}/dnn_model/data/40hc020i_0.wav")
    clean, _ = read_wav(f"{REF}/dnn_model/data/40hc020i.wav")
    n = 24000
    np.savez_compressed(f"{HERE}/wav_excerpt.npz", noisy=np.round(noisy[:n] * 32768).astype(np.int16),
                        clean=np.round(clean[:n] * 32768).astype(np.int16), fs=fs)


def graph_oracle_io():
    """o2_lstm.npz: the reference's shipped nutls_lstm.tflite executed frame by frame (oracle/tflite_graph.py) on 48
    frames of the reference's own noisy excerpt: input magnitudes, model_out, and three final history tensors."""
    from oracle.nunet_oracle import Oracle, min_max_norm
    from oracle.tflite_graph import TFLiteGraph, zero_feed
    from oracle.wavio import read_wav
    noisy, _ = read_wav(f"{REF}/dnn_model/data/40hc020i_0.wav")
    noisy = min_max_norm(noisy).astype(np.float32)
    T = 48
    seg = noisy[8000:8000 + 512 + 256 * (T - 1)]
    mags, _ = Oracle({}, ctfa_mode="frame_div32").stft(torch.from_numpy(seg)[None])
    mag = mags[0, :, 1:].numpy().astype(np.float32)
    g = TFLiteGraph(f"{REF}/dnn_model/tflite/nutls_lstm.tflite")
    feed = zero_feed(g)
    outs = []
    with torch.no_grad():
        for t in range(T):
            feed["input"] = torch.from_numpy(mag[t].reshape(1, 1, 256, 1))
            res = g.run(feed)
            outs.append(res["model_out"].reshape(256).numpy().copy())
            for k, val in res.items():
                kin = k.replace("_cur", "_prev")
                if k != "model_out" and kin in feed:
                    feed[kin] = val
    keep = {k: res[k].numpy() for k in ("msfe6_ee_cur1", "msfe4_dd2_cur3", "state_h", "msfe5_en_c")}
    np.savez_compressed(f"{HERE}/o2_lstm.npz", mag=mag, model_out=np.stack(outs), **{f"state_{k}": v for k, v in keep.items()})


def graph_oracle_ddb_io():
    """o2_ddb.npz: the shipped nutls.tflite (dilated-dense baseline) executed frame by frame on 72 frames (covers the
    32-frame history of the deepest dilation)."""
    from oracle.nunet_oracle import Oracle, min_max_norm
    from oracle.tflite_graph import TFLiteGraph, stream_frames
    from oracle.wavio import read_wav
    noisy, _ = read_wav(f"{REF}/dnn_model/data/40hc020i_0.wav")
    noisy = min_max_norm(noisy).astype(np.float32)
    T = 72
    seg = noisy[8000:8000 + 512 + 256 * (T - 1)]
    mags, _ = Oracle({}, ctfa_mode="frame_div32").stft(torch.from_numpy(seg)[None])
    mag = mags[0, :, 1:].numpy().astype(np.float32)
    est = stream_frames(TFLiteGraph(f"{REF}/dnn_model/tflite/nutls.tflite"), mag)
    np.savez_compressed(f"{HERE}/o2_ddb.npz", mag=mag, model_out=est)


def oracle_io():
    from nunet_b200.synth import synth_clips
    from nunet_b200.weights import lstm_weights_from_h5
    from oracle.nunet_oracle import Oracle
    w = lstm_weights_from_h5(f"{REF}/dnn_model/log/saved_model/nutls_lstm.h5")
    wav = synth_clips(2, 512 + 256 * 39, first_clip=0)           # 2 clips x 40 frames
    out = {"wav": wav}
    with torch.no_grad():
        for mode in ("causal_avg32", "frame_div32"):
            o = Oracle(w, ctfa_mode=mode)
            y, est = o.forward_wav(wav)
            out[f"est_{mode}"] = est.numpy()
            out[f"wav_{mode}"] = y.numpy()
    np.savez_compressed(f"{HERE}/oracle_io_lstm.npz", **out)


def reference_artifacts():
    """SHA-256 / size of the reference's shipped artefacts: the exporter tests reproduce nutls_lstm.tflite byte for byte."""
    import hashlib
    out = {}
    for name, rel in (("nutls_lstm.tflite", "dnn_model/tflite/nutls_lstm.tflite"), ("nutls.tflite", "dnn_model/tflite/nutls.tflite"),
                      ("nutls_lstm.h5", "dnn_model/log/saved_model/nutls_lstm.h5")):
        b = open(f"{REF}/{rel}", "rb").read()
        out[name] = {"sha256": hashlib.sha256(b).hexdigest(), "bytes": len(b)}
    json.dump(out, open(f"{HERE}/reference_artifacts.json", "w"), indent=1)
    from nunet_b200.tflite_export import build_skeleton
    build_skeleton()          # nunet_b200/data/nutls_lstm_skeleton.tflite.gz: the shipped graph with weights and scales zeroed


if __name__ == "__main__":
    java_windows()
    state_tables()
    wav_excerpt()
    oracle_io()
    graph_oracle_io()
    graph_oracle_ddb_io()
    reference_artifacts()
    print("fixtures written to", HERE)
