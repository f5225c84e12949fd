"""Artifact emission on the deployment side of the path: a role-named weight set -> a `.tflite` with the reference's
one-frame stateful graph and dynamic-range int8 weights, i.e. what `dnn_model/converter_proposed.py:877-912` produces
(`tf.saved_model.save(signatures=...)` -> `TFLiteConverter` with `Optimize.DEFAULT` :901) and the interpreter / the
Android app load (`interpreter_proposed.py:374`, `RTSE_NUTLS_LSTM.java`).

No TensorFlow exists here, so the graph is not re-traced: the flatbuffer STRUCTURE (3066 operators, 131-tensor signature
`nutls_lstm_sm`, tensor names, shapes, operator options) is taken from a skeleton of the reference's shipped
`tflite/nutls_lstm.tflite` in which every weight buffer and every quantisation scale has been zeroed
(`data/nutls_lstm_skeleton.tflite.gz`, made by `build_skeleton`, 0.2 MB), and the exporter fills in the weights of the set
it is given -- quantised the way the TFLite converter's weight-only ("dynamic range") pass does it: tensors that are int8 in
the reference file (>= 1024 elements) become symmetric int8, per output channel for (transpose) convolutions
(scale_c = max|w_c| / 127), per tensor for fully-connected kernels; small tensors, biases, LayerNorm and PReLU parameters
stay float32.  Because only buffer contents and scale values change, every byte offset of the flatbuffer stays valid.

The written file is a complete TFLite model: `tflite_reader.read_tflite` / `weights.lstm_weights_from_tflite` read it back,
the flatbuffer executor of the test suite (oracle/tflite_graph.py) runs it frame by frame, and `Interpreter(model_path=...)`
of this package serves it.  (The SavedModel directory of :888-893 is a TensorFlow checkpoint + protobuf graph and is not
emitted.)
"""
from __future__ import annotations

import gzip
import os
from typing import Dict, List, Tuple

import numpy as np

from .tflite_reader import _FB, read_tflite
from .weights import _classify_unit, _tflite_role_tensors

_HERE = os.path.dirname(os.path.abspath(__file__))
SKELETON_LSTM = os.path.join(_HERE, "data", "nutls_lstm_skeleton.tflite.gz")
REFERENCE_TFLITE_LSTM = "/root/reference/dnn_model/tflite/nutls_lstm.tflite"


def _from_keras_layout(w: np.ndarray, kind: str, shape: tuple) -> np.ndarray:
    """inverse of weights._to_keras_layout: the weight-set array in the tensor's TFLite layout / shape"""
    w = np.asarray(w, np.float32)
    if kind == "T":
        w = w.T
    elif kind == "conv":
        w = w.transpose(3, 0, 1, 2)
    elif kind == "tconv":
        w = w.transpose(2, 0, 1, 3)
    elif kind == "mlp":
        w = w.T.reshape(shape)
    return np.ascontiguousarray(w).reshape(shape)


def _layout(buf: bytes):
    """[(Tensor, key, kind, buffer offset, buffer bytes, scale offset, scale count)] of every weight tensor of the file."""
    import tempfile
    with tempfile.NamedTemporaryFile(suffix=".tflite", delete=False) as f:
        f.write(buf)
        tmp = f.name
    try:
        _g, by_role = _tflite_role_tensors(tmp)
    finally:
        os.unlink(tmp)
    fb = _FB(buf)
    model = fb.root()
    buffers = [fb.vector(bt, 0) for bt in fb.tables(model, 4)]
    ttables = fb.tables(fb.tables(model, 2)[0], 0)
    out = []
    for role, tensors in by_role.items():
        for t, key, kind in _classify_unit("out_conv" if role == "conv2d" else role, tensors):
            boff, bn = buffers[t.buffer]
            q = fb.table_field(ttables[t.index], 4)
            soff, sn = fb.vector(q, 2) if q else (0, 0)
            out.append((t, key, kind, boff, bn, soff, sn))
    return out


def build_skeleton(src: str = REFERENCE_TFLITE_LSTM, dst: str = SKELETON_LSTM) -> str:
    """Zero every weight buffer and quantisation scale of the reference file (build container only)."""
    buf = bytearray(open(src, "rb").read())
    for t, _key, _kind, boff, bn, soff, sn in _layout(bytes(buf)):
        buf[boff:boff + bn] = bytes(bn)
        if sn:
            buf[soff:soff + 4 * sn] = bytes(4 * sn)
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    with gzip.GzipFile(dst, "wb", compresslevel=9, mtime=0) as f:
        f.write(bytes(buf))
    return dst


def quantize_symmetric(w: np.ndarray, per_channel: bool) -> Tuple[np.ndarray, np.ndarray]:
    """TFLite weight-only quantisation (tensorflow/lite/tools/optimize quantization_utils SymmetricPerChannelQuantization /
    SymmetricQuantizeTensor): scale = max|w| / 127 (per output channel = dim 0, or per tensor), q = round-half-away(w / scale)
    clamped to [-127, 127].  Scale and quotient are evaluated in double and the scale stored as float32: with that, quantising
    the reference's float checkpoint reproduces every int8 value and scale of the reference's shipped file
    (tests/test_tflite_export.py)."""
    w = np.asarray(w, np.float32)
    if per_channel:
        mx = np.abs(w.reshape(w.shape[0], -1)).max(axis=1).astype(np.float64)
    else:
        mx = np.abs(w).max(keepdims=True).reshape(1).astype(np.float64)
    scale = mx / 127.0
    safe = np.where(scale > 0, scale, 1.0)
    x = w.astype(np.float64) / (safe.reshape((-1,) + (1,) * (w.ndim - 1)) if per_channel else safe[0])
    q = np.clip(np.sign(x) * np.floor(np.abs(x) + 0.5), -127, 127).astype(np.int8)
    return q, scale.astype(np.float32)


def export_lstm_tflite(weights: Dict[str, np.ndarray], path: str, skeleton: str = SKELETON_LSTM) -> Dict[str, int]:
    """Write `weights` (role-named NUNet-TLS-LSTM set, weights.expected_lstm_shapes) as a dynamic-range-quantised
    `.tflite` with the reference's `nutls_lstm_sm` signature.  Returns counts of int8 / float tensors written."""
    if not os.path.exists(skeleton):
        raise FileNotFoundError(f"{skeleton} missing (made by tflite_export.build_skeleton from the reference's shipped graph)")
    buf = bytearray(gzip.open(skeleton, "rb").read())
    n_q = n_f = 0
    for t, key, kind, boff, bn, soff, sn in _layout(bytes(buf)):
        if key not in weights:
            if key.endswith("/bias"):
                continue
            raise KeyError(f"weight set has no '{key}' (tensor {t.name})")
        w = _from_keras_layout(weights[key], kind, t.shape)
        if t.dtype == np.int8:
            q, scale = quantize_symmetric(w, per_channel=sn > 1)
            if q.nbytes != bn or scale.size != sn:
                raise ValueError(f"{key}: {q.nbytes} bytes / {scale.size} scales do not fit the graph's {bn} / {sn}")
            buf[boff:boff + bn] = q.tobytes()
            buf[soff:soff + 4 * sn] = scale.astype("<f4").tobytes()
            n_q += 1
        else:
            raw = w.astype("<f4").tobytes()
            if len(raw) != bn:
                raise ValueError(f"{key}: {len(raw)} bytes do not fit the graph's {bn}")
            buf[boff:boff + bn] = raw
            n_f += 1
    with open(path, "wb") as f:
        f.write(bytes(buf))
    return {"int8_tensors": n_q, "float_tensors": n_f, "bytes": len(buf)}


def ensure_skeleton() -> str:
    if not os.path.exists(SKELETON_LSTM) and os.path.exists(REFERENCE_TFLITE_LSTM):
        build_skeleton()
    return SKELETON_LSTM


def hybrid_weight_set(weights: Dict[str, np.ndarray], skeleton: str = SKELETON_LSTM) -> Dict[str, np.ndarray]:
    """The weight set of the DEPLOYED arithmetic (engine variant NUNET_VARIANT_LSTM_HYBRID): every tensor that is int8 in the
    reference's graph is quantised exactly as the exporter writes it, and goes into the set three times -- `<key>` dequantised
    (q * scale; what float operators such as the transpose convolution see after DEQUANTIZE), `<key>_q` the int8 values in the
    TFLite layout ([Cout,kh,kw,Cin] / [out,in]; stored as exact floats) and `<key>_scale` (per output channel for convolutions,
    one value for fully-connected kernels).  Float tensors are passed through.  Quantising the reference's float checkpoint
    this way gives precisely the int8 tensors of the reference's shipped nutls_lstm.tflite."""
    if not os.path.exists(skeleton):
        raise FileNotFoundError(f"{skeleton} missing (made by tflite_export.build_skeleton from the reference's shipped graph)")
    from .weights import _to_keras_layout
    out: Dict[str, np.ndarray] = {k: np.asarray(v, np.float32) for k, v in weights.items()}
    for t, key, kind, _boff, _bn, _soff, sn in _layout(gzip.open(skeleton, "rb").read()):
        if t.dtype != np.int8 or key not in weights:
            continue
        w = _from_keras_layout(weights[key], kind, t.shape)
        q, scale = quantize_symmetric(w, per_channel=sn > 1)
        deq = q.astype(np.float32) * (scale.reshape((-1,) + (1,) * (q.ndim - 1)) if sn > 1 else scale[0])
        out[key] = _to_keras_layout(deq.astype(np.float32), kind)
        out[key + "_q"] = q.astype(np.float32)
        out[key + "_scale"] = scale.astype(np.float32)
    return out
