"""Role-named weight sets for the NUNet-TLS path and the packed blob the C-ABI consumes.

A *weight set* is `{"<layer>/<var>": float32 ndarray}` in the reference's own (Keras) tensor
layouts, where `<layer>` is the layer's *role* in `models/proposed.py:284-625`
(`msfe6_en_conv1`, `msfe4_de2_spconv3`, `lstm`, `out_conv`, ...) and `<var>` one of
`kernel, bias, gamma, beta, alpha` (conv units), `kernel0, bias0, kernel1, bias1` (the two 1x1
layers of a CTFA time/frequency attention MLP), `kernel, recurrent_kernel, bias` (LSTM).

The `.h5` groups weights by the *training model's* layer names, and six decoder Dense layers are
mis-named there (`proposed.py:47-63`); Keras loads them topologically (`converter_proposed.py:13`),
so the role -> group remap below is part of the format.
"""
from __future__ import annotations

import re
import struct
from typing import Dict, Optional

import numpy as np

from .h5_reader import read_h5

MAGIC = b"NUNETW01"
VARIANT_LSTM = 0
VARIANT_DDB = 1
VARIANT_LSTM_HYBRID = 2     # NUNet-TLS-LSTM with the deployed int8 dynamic-range arithmetic (tflite_export.hybrid_weight_set)
_ENTRY = struct.Struct("<64sI4IQ4x")  # name, ndim, dims[4], offset (floats), pad -> 96 bytes
assert _ENTRY.size == 96

# role (Dense that follows <role>_lstm) -> h5 group that actually holds it (proposed.py:47-63)
DENSE_ROLE_TO_H5 = {
    "msfe3_de_dense": "msfe6_de_dense",
    "msfe4_de_dense": "msfe5_de_dense",
    "msfe4_de2_dense": "msfe4_de_dense",
    "msfe4_de3_dense": "msfe4_de2_dense",
    "msfe5_de_dense": "msfe4_de3_dense",
    "msfe6_de_dense": "msfe3_de_dense",
}
_H5_TO_DENSE_ROLE = {v: k for k, v in DENSE_ROLE_TO_H5.items()}

# (block prefix, F0, depth n) in network order; encoder side then decoder side (proposed.py:297-613)
ENC_BLOCKS = [("msfe6_en", 256, 6), ("msfe5_en", 128, 5), ("msfe4_en", 64, 4),
              ("msfe4_en2", 32, 4), ("msfe4_en3", 16, 4), ("msfe3_en", 8, 3)]
DEC_BLOCKS = [("msfe3_de", 8, 3), ("msfe4_de", 16, 4), ("msfe4_de2", 32, 4),
              ("msfe4_de3", 64, 4), ("msfe5_de", 128, 5), ("msfe6_de", 256, 6)]
DOWN_NAMES = ["msfe6_down_sampling", "msfe5_down_sampling", "msfe4_down_sampling",
              "msfe4_down_sampling2", "msfe4_down_sampling3", "msfe3_down_sampling"]
UP_NAMES = ["msfe3_upsampling", "msfe4_upsampling", "msfe4_upsampling2",
            "msfe4_upsampling3", "msfe5_upsampling", "msfe6_upsampling"]


def _suffix(name: str) -> int:
    m = re.search(r"_(\d+)$", name)
    return int(m.group(1)) if m else 0


def lstm_weights_from_h5(path: str, trace: Optional[Dict[str, str]] = None) -> Dict[str, np.ndarray]:
    """Role-named float32 weight set of NUNet-TLS-LSTM from the reference `.h5`.  `trace`, when given, receives
    role key -> dataset path (what `keras_export` needs to write the file back)."""
    raw = read_h5(path)
    groups: Dict[str, Dict[str, Dict[str, np.ndarray]]] = {}
    paths: Dict[tuple, str] = {}
    for key, arr in raw.items():
        parts = key.strip("/").split("/")
        group, sub, var = parts[0], parts[-2], parts[-1].split(":")[0]
        groups.setdefault(group, {}).setdefault(sub, {})[var] = arr.astype(np.float32)
        paths[(group, sub, var)] = key

    out: Dict[str, np.ndarray] = {}

    def put(role_key: str, group: str, sub: str, var: str, value: np.ndarray) -> None:
        out[role_key] = value
        if trace is not None:
            trace[role_key] = paths[(group, sub, var)]

    for group, subs in groups.items():
        role = _H5_TO_DENSE_ROLE.get(group, group)
        if group == "conv2d":
            role = "out_conv"
        mlp = 0
        for sub in sorted(subs, key=lambda s: (_suffix(s), s)):
            v = subs[sub]
            if sub.startswith("layer_normalization"):
                put(f"{role}/gamma", group, sub, "gamma", v["gamma"])
                put(f"{role}/beta", group, sub, "beta", v["beta"])
            elif sub.startswith("p_re_lu"):
                put(f"{role}/alpha", group, sub, "alpha", v["alpha"].reshape(1))
            elif sub.startswith("lstm_cell"):
                for var in ("kernel", "recurrent_kernel", "bias"):
                    put(f"{role}/{var}", group, sub, var, v[var])
            elif role.endswith("_ta") or role.endswith("_fa"):
                put(f"{role}/kernel{mlp}", group, sub, "kernel", v["kernel"].reshape(v["kernel"].shape[-2:]))
                put(f"{role}/bias{mlp}", group, sub, "bias", v["bias"])
                mlp += 1
            else:  # Conv2D / Conv2DTranspose / Dense
                put(f"{role}/kernel", group, sub, "kernel", v["kernel"])
                put(f"{role}/bias", group, sub, "bias", v["bias"])
    return out


def _tflite_role_tensors(path: str):
    """role -> list of constant tensors of a shipped .tflite whose name starts with the Keras layer name
    (`<layer>/<sub-layer>/<op>`).  Returns (graph, {role: [Tensor]})."""
    from .tflite_reader import read_tflite
    g = read_tflite(path)
    by_role: Dict[str, list] = {}
    for t in g.tensors:
        if t.data is None or t.data.size == 0 or t.dtype not in (np.float32, np.int8):
            continue
        first = t.name.split(";")[0]
        if "/" not in first:
            continue
        by_role.setdefault(first.split("/")[0], []).append(t)
    return g, by_role


def _classify_unit(role: str, tensors):
    """[(tensor, weight-set key, layout kind)] for the constants attributed to layer `role` of a shipped .tflite.
    kind: 'asis'; 'vec1' (scalar alpha); 'T' ([out,in] -> (in,out)); 'conv' ([Cout,kh,kw,Cin] -> (kh,kw,Cin,Cout));
    'tconv' ([Cout,kh,kw,Cin] -> (kh,kw,Cout,Cin)); 'mlp' (1x1 conv [Cout,1,1,Cin] -> (Cin,Cout)).
    Shared by the importer below and the exporter (tflite_export.py), which applies the inverse."""
    res, convs, biases = [], [], {}
    for t in tensors:
        first = t.name.split(";")[0]
        sub = first.split("/")[1]
        leaf = first.split("/", 2)[2] if first.count("/") >= 2 else ""
        nd = len(t.shape)
        if sub.startswith("layer_normalization"):
            if leaf == "batchnorm/mul/ReadVariableOp":
                res.append((t, f"{role}/gamma", "asis"))
            elif leaf == "batchnorm/ReadVariableOp":
                res.append((t, f"{role}/beta", "asis"))
        elif sub.startswith("p_re_lu"):
            res.append((t, f"{role}/alpha", "vec1"))
        elif sub.startswith("lstm_cell"):
            if nd == 2 and tuple(t.shape) == (84, 21):      # leaf names vary (MatMul_1, MatMul_11): go by shape
                res.append((t, f"{role}/recurrent_kernel", "T"))
            elif nd == 2:
                res.append((t, f"{role}/kernel", "T"))
            elif leaf.startswith("BiasAdd"):
                res.append((t, f"{role}/bias", "asis"))
        elif sub == "Tensordot":                                # Dense applied to [T, 21]
            res.append((t, f"{role}/kernel", "T"))
        elif sub == "BiasAdd":
            res.append((t, f"{role}/bias", "asis"))
        elif sub.startswith("conv2d_transpose"):
            res.append((t, f"{role}/kernel", "tconv") if nd == 4 else (t, f"{role}/bias", "asis"))
        elif sub.startswith("conv") or sub == "Conv2D":
            if nd == 4:
                convs.append((_suffix(sub), t))
            elif leaf.startswith("BiasAdd"):
                biases[_suffix(sub)] = t
    convs.sort(key=lambda c: c[0])
    if role.endswith("_ta") or role.endswith("_fa"):
        for i, (suf, t) in enumerate(convs):
            res.append((t, f"{role}/kernel{i}", "mlp"))
            res.append((biases[suf], f"{role}/bias{i}", "asis"))
    elif convs:
        suf, t = convs[0]
        res.append((t, f"{role}/kernel", "conv"))
        if suf in biases:
            res.append((biases[suf], f"{role}/bias", "asis"))
    return res


def _to_keras_layout(w: np.ndarray, kind: str) -> np.ndarray:
    if kind == "vec1":
        return w.reshape(1)
    if kind == "T":
        return np.ascontiguousarray(w.T)
    if kind == "conv":
        return np.ascontiguousarray(w.transpose(1, 2, 3, 0))
    if kind == "tconv":
        return np.ascontiguousarray(w.transpose(1, 2, 0, 3))
    if kind == "mlp":
        k = np.ascontiguousarray(w.transpose(1, 2, 3, 0))
        return k.reshape(k.shape[-2:])
    return w


def _unit_from_tflite(role: str, tensors, out: Dict[str, np.ndarray]) -> None:
    """Fill the `<role>/...` entries of a weight set from the constants attributed to that layer, converting
    TFLite layouts to the Keras ones: conv [Cout,kh,kw,Cin] -> (kh,kw,Cin,Cout); transpose conv
    [Cout,kh,kw,Cin] -> (kh,kw,Cout,Cin); FC [out,in] -> (in,out)."""
    for t, key, kind in _classify_unit(role, tensors):
        out[key] = _to_keras_layout(t.dequantized(), kind)


def lstm_weights_from_tflite(path: str) -> Dict[str, np.ndarray]:
    """Role-named float32 weight set of NUNet-TLS-LSTM from the shipped `nutls_lstm.tflite`: int8 tensors are
    dequantised (q * scale, per output channel for conv, per tensor for FC).  Used to run the source restatement
    and the flatbuffer executor with IDENTICAL weights."""
    _g, by_role = _tflite_role_tensors(path)
    out: Dict[str, np.ndarray] = {}
    for role, tensors in by_role.items():
        _unit_from_tflite("out_conv" if role == "conv2d" else role, tensors, out)
    for un in UP_NAMES:                                         # no bias tensor in the graph when it is all zeros
        out.setdefault(f"{un}/bias", np.zeros(128, np.float32))
    return out


DDB_DILATIONS = (1, 2, 4, 8, 16, 32)


def ddb_roles():
    """(role prefix, channels C) of the 13 dilated-dense bottlenecks in network order (nunet_tls.py:383-410, 678-700)."""
    out = [(f"{name}_ddb", 32) for name, _f0, _n in ENC_BLOCKS]
    out.append(("ddb", 64))
    out += [(f"{name}_ddb", 32) for name, _f0, _n in DEC_BLOCKS]
    return out


def ddb_weights_from_tflite(path: str) -> Dict[str, np.ndarray]:
    """Role-named float32 weight set of the NUNet-TLS baseline (dilated-dense bottleneck) from the shipped
    `nutls.tflite`, its ONLY weight source.  Per DDB layer k: `kernel0/bias0` = the grouped dilated
    (2,3) conv in Keras layout (2,3,k,h), `kernel1/bias1` = the 1x1 conv (h,h), `gamma/beta/alpha`.
    The grouped kernels are anonymous tensors (`Conv2DNN`) in the flatbuffer and are attributed through the graph:
    by the conv's own bias tensor (dilation 1) or by the bias ADD that follows BATCH_TO_SPACE_ND (dilation >= 2,
    lowered to SPACE_TO_BATCH / CONV / BATCH_TO_SPACE).  The three MSFE4 encoder down-sampling convs share one weight
    tensor in the shipped file and are aliased here."""
    g, by_role = _tflite_role_tensors(path)
    out: Dict[str, np.ndarray] = {}
    ddb_re = re.compile(r"^(.*ddb)_(\d)$")
    for role, tensors in by_role.items():
        if role.startswith("Conv2D"):
            continue
        m = ddb_re.match(role)
        if not m:
            _unit_from_tflite("out_conv" if role == "conv2d" else role, tensors, out)
            continue
        biases = []
        for t in tensors:
            first = t.name.split(";")[0]
            sub = first.split("/")[1]
            leaf = first.split("/", 2)[2] if first.count("/") >= 2 else ""
            w = t.dequantized()
            if sub.startswith("layer_normalization"):
                if leaf == "batchnorm/mul/ReadVariableOp":
                    out[f"{role}/gamma"] = w
                elif leaf == "batchnorm/ReadVariableOp":
                    out[f"{role}/beta"] = w
            elif sub.startswith("p_re_lu"):
                out[f"{role}/alpha"] = w.reshape(1)
            elif sub.startswith("conv2d") and w.ndim == 4:
                out[f"{role}/kernel1"] = np.ascontiguousarray(w.transpose(1, 2, 3, 0)).reshape(w.shape[3], w.shape[0])
            elif sub.startswith("conv2d") and leaf.startswith("BiasAdd"):
                biases.append((_suffix(sub), w))
        biases.sort(key=lambda b: b[0])
        out[f"{role}/bias0"], out[f"{role}/bias1"] = biases[0][1], biases[1][1]
    # anonymous grouped-conv kernels
    producer = {}
    consumers: Dict[int, list] = {}
    for op in g.operators:
        for o in op.outputs:
            producer[o] = op
        for i in op.inputs:
            if i >= 0:
                consumers.setdefault(i, []).append(op)

    def role_of(tname: str):
        first = tname.split(";")[0]
        return first.split("/")[0] if "/" in first and ddb_re.match(first.split("/")[0]) else None

    for op in g.operators:
        if op.op != "CONV_2D" or not g.tensors[op.inputs[1]].name.startswith("Conv2D"):
            continue
        role = role_of(g.tensors[op.inputs[2]].name) if len(op.inputs) > 2 else None
        if role is None:      # dilated lowering: CONV_2D -> BATCH_TO_SPACE_ND -> ADD(bias)
            nxt = consumers[op.outputs[0]][0]
            assert nxt.op == "BATCH_TO_SPACE_ND", nxt.op
            add = consumers[nxt.outputs[0]][0]
            assert add.op == "ADD", add.op
            for i in add.inputs:
                role = role or role_of(g.tensors[i].name)
        assert role is not None, g.tensors[op.inputs[1]].name
        w = g.tensors[op.inputs[1]].dequantized()                 # [h, 2, 3, k]
        out[f"{role}/kernel0"] = np.ascontiguousarray(w.transpose(1, 2, 3, 0))
    for alias in ("msfe4_down_sampling2", "msfe4_down_sampling3"):
        for var in ("kernel", "bias"):
            out.setdefault(f"{alias}/{var}", out[f"msfe4_down_sampling/{var}"].copy())
    for un in UP_NAMES:
        out.setdefault(f"{un}/bias", np.zeros(128, np.float32))
    return out


def expected_ddb_shapes() -> Dict[str, tuple]:
    """Shape table of the dilated-dense variant: the LSTM table with every LSTM + Dense replaced by a DDB."""
    s = {k: v for k, v in expected_lstm_shapes().items()
         if not (k.split("/")[0].endswith("_lstm") or k.split("/")[0].endswith("_dense") or k.split("/")[0] in ("lstm", "dense"))}
    for role, C in ddb_roles():
        h = C // 2
        s[f"{role}_in/kernel"], s[f"{role}_in/bias"], s[f"{role}_in/alpha"] = (2, 3, C, h), (h,), (1,)
        for k in range(1, 7):
            r = f"{role}_{k}"
            s[f"{r}/kernel0"], s[f"{r}/bias0"] = (2, 3, k, h), (h,)
            s[f"{r}/kernel1"], s[f"{r}/bias1"] = (h, h), (h,)
            s[f"{r}/gamma"] = s[f"{r}/beta"] = (h,)
            s[f"{r}/alpha"] = (1,)
        s[f"{role}_out/kernel"], s[f"{role}_out/bias"], s[f"{role}_out/alpha"] = (2, 3, h, C), (C,), (1,)
    return s


def expected_lstm_shapes() -> Dict[str, tuple]:
    """Shape table of the LSTM variant, derived from the topology, used to validate a set."""
    s: Dict[str, tuple] = {}

    def unit(name, kshape, ln):
        s[f"{name}/kernel"] = kshape
        s[f"{name}/bias"] = (kshape[-1],)
        if ln:
            s[f"{name}/gamma"] = s[f"{name}/beta"] = (ln,)
            s[f"{name}/alpha"] = (1,)

    unit("input_layer", (1, 1, 1, 64), 64)
    for side, blocks in (("en", ENC_BLOCKS), ("de", DEC_BLOCKS)):
        for name, f0, n in blocks:
            unit(f"{name}_in", (1, 1, 64 if side == "en" else 128, 64), 64)
            for k in range(1, n + 1):
                if side == "en":
                    cin = 64 if k == 1 else 32
                else:
                    cin = 128 if k == 1 else 64
                unit(f"{name}_conv{k}", (2, 3, cin, 32), 32)
            d = (f0 >> n) * 32
            s[f"{name}_lstm/kernel"] = (d, 84)
            s[f"{name}_lstm/recurrent_kernel"] = (21, 84)
            s[f"{name}_lstm/bias"] = (84,)
            s[f"{name}_dense/kernel"] = (21, d)
            s[f"{name}_dense/bias"] = (d,)
            for k in range(1, n + 1):
                co = 64 if k == n else 32
                unit(f"{name}_spconv{k}", (2, 3, 64, 2 * co), co)
            for att in ("ta", "fa"):
                s[f"{name}_{att}/kernel0"] = (64, 16)
                s[f"{name}_{att}/bias0"] = (16,)
                s[f"{name}_{att}/kernel1"] = (16, 64)
                s[f"{name}_{att}/bias1"] = (64,)
    for dn in DOWN_NAMES:
        unit(dn, (1, 3, 64, 64), 0)
    for un in UP_NAMES:
        s[f"{un}/kernel"] = (1, 3, 128, 128)
        s[f"{un}/bias"] = (128,)
    s["lstm/kernel"], s["lstm/recurrent_kernel"], s["lstm/bias"] = (256, 84), (21, 84), (84,)
    s["dense/kernel"], s["dense/bias"] = (21, 256), (256,)
    unit("out_conv", (1, 1, 64, 1), 0)
    return s


def validate(weights: Dict[str, np.ndarray], shapes: Dict[str, tuple]) -> None:
    missing = sorted(set(shapes) - set(weights))
    extra = sorted(set(weights) - set(shapes))
    if missing or extra:
        raise ValueError(f"weight set mismatch: missing {missing[:5]} extra {extra[:5]}")
    for k, shp in shapes.items():
        if tuple(weights[k].shape) != tuple(shp):
            raise ValueError(f"{k}: shape {weights[k].shape} != expected {shp}")


def pack_blob(weights: Dict[str, np.ndarray], variant: int = VARIANT_LSTM) -> bytes:
    """Serialise a weight set into the blob `nunet_create` takes (include/nunet_b200.h)."""
    names = sorted(weights)
    head = bytearray(MAGIC + struct.pack("<II", len(names), variant))
    data = []
    off = 0
    for n in names:
        a = np.ascontiguousarray(weights[n], dtype="<f4")
        if a.ndim > 4 or len(n.encode()) > 63:
            raise ValueError(f"cannot pack {n} {a.shape}")
        dims = list(a.shape) + [0] * (4 - a.ndim)
        head += _ENTRY.pack(n.encode(), a.ndim, *dims, off)
        data.append(a.tobytes())
        off += a.size
    return bytes(head) + b"".join(data)


def unpack_blob(blob: bytes):
    if blob[:8] != MAGIC:
        raise ValueError("bad blob magic")
    n, variant = struct.unpack_from("<II", blob, 8)
    base = 16 + n * _ENTRY.size
    out = {}
    for i in range(n):
        name, ndim, d0, d1, d2, d3, off = _ENTRY.unpack_from(blob, 16 + i * _ENTRY.size)
        shape = (d0, d1, d2, d3)[:ndim]
        cnt = int(np.prod(shape)) if shape else 1
        out[name.rstrip(b"\0").decode()] = np.frombuffer(blob, "<f4", cnt, base + 4 * off).reshape(shape).copy()
    return out, variant


# ---------------------------------------------------------------------------------------------------
# Default weight artefact.  The reference checkout (and therefore the .h5) does not exist on the GPU box, so
# `__graft_entry__.build()` converts it once into nunet_b200/data/nutls_lstm.nunetw.  The three blobs under data/ are
# COMMITTED: they are data extracted from the reference's own .h5 / .tflite by the code above, and the product needs
# them at run time on boxes without the reference checkout.  Nothing here fabricates weights silently: random weights
# must be asked for by name.
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_BLOB = os.path.join(_HERE, "data", "nutls_lstm.nunetw")
REFERENCE_H5 = "/root/reference/dnn_model/log/saved_model/nutls_lstm.h5"
# the same architecture with the weights of the reference's SHIPPED graph (int8 tensors dequantised): lets the engine
# and the source restatement be compared with the flatbuffer executor on identical weights
TFLITE_LSTM_BLOB = os.path.join(_HERE, "data", "nutls_lstm_tflite.nunetw")
REFERENCE_TFLITE_LSTM = "/root/reference/dnn_model/tflite/nutls_lstm.tflite"


def ensure_default_blob() -> str:
    """Create DEFAULT_BLOB from the reference .h5 when the reference checkout is present."""
    if not os.path.exists(DEFAULT_BLOB) and os.path.exists(REFERENCE_H5):
        w = lstm_weights_from_h5(REFERENCE_H5)
        validate(w, expected_lstm_shapes())
        os.makedirs(os.path.dirname(DEFAULT_BLOB), exist_ok=True)
        with open(DEFAULT_BLOB, "wb") as f:
            f.write(pack_blob(w))
    return DEFAULT_BLOB


DDB_BLOB = os.path.join(_HERE, "data", "nutls_ddb.nunetw")
REFERENCE_TFLITE_DDB = "/root/reference/dnn_model/tflite/nutls.tflite"


def ensure_ddb_blob() -> str:
    """The dilated-dense variant's only weights: the shipped nutls.tflite, dequantised."""
    if not os.path.exists(DDB_BLOB) and os.path.exists(REFERENCE_TFLITE_DDB):
        w = ddb_weights_from_tflite(REFERENCE_TFLITE_DDB)
        validate(w, expected_ddb_shapes())
        os.makedirs(os.path.dirname(DDB_BLOB), exist_ok=True)
        with open(DDB_BLOB, "wb") as f:
            f.write(pack_blob(w, VARIANT_DDB))
    return DDB_BLOB


def load_ddb_weights() -> Dict[str, np.ndarray]:
    path = ensure_ddb_blob()
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing and {REFERENCE_TFLITE_DDB} not available")
    with open(path, "rb") as f:
        w, variant = unpack_blob(f.read())
    if variant != VARIANT_DDB:
        raise ValueError("not a dilated-dense blob")
    return w


def ensure_tflite_lstm_blob() -> str:
    if not os.path.exists(TFLITE_LSTM_BLOB) and os.path.exists(REFERENCE_TFLITE_LSTM):
        w = lstm_weights_from_tflite(REFERENCE_TFLITE_LSTM)
        validate(w, expected_lstm_shapes())
        os.makedirs(os.path.dirname(TFLITE_LSTM_BLOB), exist_ok=True)
        with open(TFLITE_LSTM_BLOB, "wb") as f:
            f.write(pack_blob(w))
    return TFLITE_LSTM_BLOB


def load_tflite_lstm_weights() -> Dict[str, np.ndarray]:
    path = ensure_tflite_lstm_blob()
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing and {REFERENCE_TFLITE_LSTM} not available")
    with open(path, "rb") as f:
        return unpack_blob(f.read())[0]


def load_default_weights() -> Dict[str, np.ndarray]:
    path = ensure_default_blob()
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing and {REFERENCE_H5} not available: run __graft_entry__.build() "
                                "where the reference checkout is mounted")
    with open(path, "rb") as f:
        w, variant = unpack_blob(f.read())
    if variant != VARIANT_LSTM:
        raise ValueError("default blob is not the LSTM variant")
    return w


def random_ddb_weights(seed: int = 0) -> Dict[str, np.ndarray]:
    """Seeded random-init weight set of the dilated-dense variant (plumbing / parity tests with O(1) activations)."""
    rng = np.random.default_rng(seed)
    out: Dict[str, np.ndarray] = {}
    for k, shp in expected_ddb_shapes().items():
        var = k.split("/")[1]
        if var == "gamma":
            a = 1.0 + 0.1 * rng.standard_normal(shp)
        elif var == "alpha":
            a = np.full(shp, 0.25)
        elif var.startswith("bias") or var == "beta":
            a = 0.05 * rng.standard_normal(shp)
        else:
            fan_in = int(np.prod(shp[:-1]))
            a = rng.standard_normal(shp) / np.sqrt(max(fan_in, 1))
        out[k] = a.astype(np.float32)
    return out


def random_lstm_weights(seed: int = 0) -> Dict[str, np.ndarray]:
    """Seeded random-init weight set of the NUNet-TLS-LSTM architecture (Glorot-like scale); for shape /
    plumbing tests and as an explicitly-labelled stand-in when the trained checkpoint is unavailable."""
    rng = np.random.default_rng(seed)
    out: Dict[str, np.ndarray] = {}
    for k, shp in expected_lstm_shapes().items():
        var = k.split("/")[1]
        if var == "gamma":
            a = 1.0 + 0.1 * rng.standard_normal(shp)
        elif var == "alpha":
            a = np.full(shp, 0.25)
        elif var.startswith("bias") or var == "beta":
            a = 0.05 * rng.standard_normal(shp)
        else:
            fan_in = int(np.prod(shp[:-1]))
            a = rng.standard_normal(shp) / np.sqrt(max(fan_in, 1))
        out[k] = a.astype(np.float32)
    return out
