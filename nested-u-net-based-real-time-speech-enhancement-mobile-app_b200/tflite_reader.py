"""Dependency-free reader for the subset of the TFLite FlatBuffer format (`TFL3`, schema v3) that the reference's
shipped graphs use (`dnn_model/tflite/nutls_lstm.tflite`, `nutls.tflite`; neither `flatbuffers` nor `tflite_runtime`
is installable here).  Field indices follow the public schema (tensorflow/lite/schema/schema.fbs); see SURVEY.md
Appendix A.2 for the list that matters.

Product use: the only weight source of the dilated-dense (DDB) variant is `nutls.tflite` (weights.py dequantises
its int8 tensors).  Test use: `oracle/tflite_graph.py` executes the graphs themselves.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

BUILTIN_NAMES = {
    0: "ADD", 1: "AVERAGE_POOL_2D", 2: "CONCATENATION", 3: "CONV_2D", 6: "DEQUANTIZE", 9: "FULLY_CONNECTED",
    14: "LOGISTIC", 18: "MUL", 22: "RESHAPE", 28: "TANH", 34: "PAD", 36: "GATHER", 37: "BATCH_TO_SPACE_ND",
    38: "SPACE_TO_BATCH_ND", 39: "TRANSPOSE", 40: "MEAN", 41: "SUB", 45: "STRIDED_SLICE", 49: "SPLIT",
    54: "PRELU", 67: "TRANSPOSE_CONV", 70: "EXPAND_DIMS", 76: "RSQRT", 77: "SHAPE", 81: "REDUCE_PROD", 83: "PACK",
    88: "UNPACK", 99: "SQUARED_DIFFERENCE", 19: "RELU", 25: "SOFTMAX", 43: "SQUEEZE", 53: "CAST", 102: "SPLIT_V",
    4: "DEPTHWISE_CONV_2D", 114: "QUANTIZE",
}
TYPE_NP = {0: np.float32, 1: np.float16, 2: np.int32, 3: np.uint8, 4: np.int64, 6: np.bool_, 7: np.int16, 9: np.int8}


class _FB:
    """Minimal FlatBuffer navigation (little endian, 32-bit offsets, vtables)."""

    def __init__(self, buf: bytes):
        self.b = buf

    def u8(self, o): return self.b[o]
    def i8(self, o): return struct.unpack_from("<b", self.b, o)[0]
    def u16(self, o): return struct.unpack_from("<H", self.b, o)[0]
    def i32(self, o): return struct.unpack_from("<i", self.b, o)[0]
    def u32(self, o): return struct.unpack_from("<I", self.b, o)[0]

    def root(self) -> int:
        return self.u32(0)

    def field(self, table: int, idx: int) -> int:
        """Absolute offset of field `idx` of the table at `table`, or 0 when absent."""
        vt = table - self.i32(table)
        vsize = self.u16(vt)
        slot = 4 + 2 * idx
        if slot >= vsize:
            return 0
        off = self.u16(vt + slot)
        return table + off if off else 0

    def indirect(self, o: int) -> int:
        return o + self.u32(o)

    def table_field(self, table: int, idx: int) -> int:
        f = self.field(table, idx)
        return self.indirect(f) if f else 0

    def vector(self, table: int, idx: int):
        """(start offset of elements, length) of a vector field, or (0, 0)."""
        f = self.field(table, idx)
        if not f:
            return 0, 0
        v = self.indirect(f)
        return v + 4, self.u32(v)

    def string(self, table: int, idx: int) -> str:
        start, n = self.vector(table, idx)
        return self.b[start:start + n].decode("utf-8", "replace") if start else ""

    def scalar(self, table: int, idx: int, fmt: str, default=0):
        f = self.field(table, idx)
        return struct.unpack_from("<" + fmt, self.b, f)[0] if f else default

    def np_vector(self, table: int, idx: int, dtype) -> np.ndarray:
        start, n = self.vector(table, idx)
        if not start:
            return np.zeros(0, dtype)
        return np.frombuffer(self.b, dtype=dtype, count=n, offset=start).copy()

    def tables(self, table: int, idx: int) -> List[int]:
        start, n = self.vector(table, idx)
        return [self.indirect(start + 4 * i) for i in range(n)]


@dataclass
class Tensor:
    index: int
    name: str
    shape: tuple
    dtype: type
    buffer: int
    scale: np.ndarray
    zero_point: np.ndarray
    quantized_dimension: int
    data: Optional[np.ndarray] = None      # constant contents (raw dtype), None for activations

    def dequantized(self) -> np.ndarray:
        """float32 contents of a constant: q * scale (per-tensor or along `quantized_dimension`), zero-point applied."""
        if self.data is None:
            raise ValueError(f"tensor {self.name} is not a constant")
        if self.dtype == np.float32 or self.scale.size == 0:
            return self.data.astype(np.float32)
        q = self.data.astype(np.float32)
        zp = self.zero_point.astype(np.float32) if self.zero_point.size else np.zeros(1, np.float32)
        if self.scale.size == 1:
            return (q - zp[0]) * self.scale[0]
        shape = [1] * q.ndim
        shape[self.quantized_dimension] = -1
        zpv = zp.reshape(shape) if zp.size == self.scale.size else zp[0]
        return (q - zpv) * self.scale.reshape(shape)


@dataclass
class Operator:
    index: int
    op: str
    inputs: List[int]
    outputs: List[int]
    options: Dict[str, object] = field(default_factory=dict)


@dataclass
class Signature:
    key: str
    inputs: Dict[str, int]
    outputs: Dict[str, int]


@dataclass
class Graph:
    tensors: List[Tensor]
    operators: List[Operator]
    inputs: List[int]
    outputs: List[int]
    signatures: List[Signature]
    description: str = ""


def _options(fb: _FB, op_table: int, op: str) -> Dict[str, object]:
    t = fb.table_field(op_table, 4)   # builtin_options (union value)
    if not t:
        return {}
    s = fb.scalar
    if op == "CONV_2D":
        return {"padding": s(t, 0, "b"), "stride_w": s(t, 1, "i", 1), "stride_h": s(t, 2, "i", 1), "act": s(t, 3, "b"),
                "dilation_w": s(t, 4, "i", 1), "dilation_h": s(t, 5, "i", 1)}
    if op == "DEPTHWISE_CONV_2D":
        return {"padding": s(t, 0, "b"), "stride_w": s(t, 1, "i", 1), "stride_h": s(t, 2, "i", 1), "depth_multiplier": s(t, 3, "i", 1),
                "act": s(t, 4, "b"), "dilation_w": s(t, 5, "i", 1), "dilation_h": s(t, 6, "i", 1)}
    if op == "AVERAGE_POOL_2D":
        return {"padding": s(t, 0, "b"), "stride_w": s(t, 1, "i", 1), "stride_h": s(t, 2, "i", 1), "filter_w": s(t, 3, "i", 1),
                "filter_h": s(t, 4, "i", 1), "act": s(t, 5, "b")}
    if op == "TRANSPOSE_CONV":
        return {"padding": s(t, 0, "b"), "stride_w": s(t, 1, "i", 1), "stride_h": s(t, 2, "i", 1)}
    if op == "CONCATENATION":
        return {"axis": s(t, 0, "i"), "act": s(t, 1, "b")}
    if op in ("MEAN", "REDUCE_PROD"):
        return {"keep_dims": bool(s(t, 0, "b"))}
    if op == "STRIDED_SLICE":
        return {"begin_mask": s(t, 0, "i"), "end_mask": s(t, 1, "i"), "ellipsis_mask": s(t, 2, "i"),
                "new_axis_mask": s(t, 3, "i"), "shrink_axis_mask": s(t, 4, "i")}
    if op == "SPLIT":
        return {"num_splits": s(t, 0, "i")}
    if op == "PACK":
        return {"values_count": s(t, 0, "i"), "axis": s(t, 1, "i")}
    if op == "UNPACK":
        return {"num": s(t, 0, "i"), "axis": s(t, 1, "i")}
    if op == "GATHER":
        return {"axis": s(t, 0, "i"), "batch_dims": s(t, 1, "i")}
    if op == "FULLY_CONNECTED":
        return {"act": s(t, 0, "b"), "weights_format": s(t, 1, "b"), "keep_num_dims": bool(s(t, 2, "b")),
                "asymmetric_quantize_inputs": bool(s(t, 3, "b"))}
    if op == "RESHAPE":
        return {"new_shape": fb.np_vector(t, 0, np.int32)}
    if op in ("ADD", "SUB", "MUL"):
        return {"act": s(t, 0, "b")}
    if op == "SQUEEZE":
        return {"squeeze_dims": fb.np_vector(t, 0, np.int32)}
    return {}


def read_tflite(path: str) -> Graph:
    buf = open(path, "rb").read()
    if buf[4:8] != b"TFL3":
        raise ValueError(f"{path}: not a TFL3 flatbuffer")
    fb = _FB(buf)
    model = fb.root()
    opcodes = []
    for oc in fb.tables(model, 1):
        code = fb.scalar(oc, 3, "i", 0)
        if code == 0:
            code = fb.scalar(oc, 0, "b", 0)    # deprecated_builtin_code (< 127)
        opcodes.append(BUILTIN_NAMES.get(code, f"OP_{code}"))
    buffers = []
    for bt in fb.tables(model, 4):
        start, n = fb.vector(bt, 0)
        buffers.append((start, n))
    sub = fb.tables(model, 2)[0]
    tensors: List[Tensor] = []
    for i, tt in enumerate(fb.tables(sub, 0)):
        shape = tuple(int(v) for v in fb.np_vector(tt, 0, np.int32))
        ttype = fb.scalar(tt, 1, "b", 0)
        dtype = TYPE_NP.get(ttype, np.float32)
        bidx = fb.scalar(tt, 2, "I", 0)
        name = fb.string(tt, 3)
        q = fb.table_field(tt, 4)
        scale = fb.np_vector(q, 2, np.float32) if q else np.zeros(0, np.float32)
        zp = fb.np_vector(q, 3, np.int64) if q else np.zeros(0, np.int64)
        qdim = fb.scalar(q, 6, "i", 0) if q else 0
        data = None
        start, n = buffers[bidx] if bidx < len(buffers) else (0, 0)
        if start and n:
            data = np.frombuffer(buf, dtype=dtype, count=n // np.dtype(dtype).itemsize, offset=start).reshape(shape).copy()
        tensors.append(Tensor(i, name, shape, dtype, bidx, scale, zp, qdim, data))
    operators: List[Operator] = []
    for i, ot in enumerate(fb.tables(sub, 3)):
        op = opcodes[fb.scalar(ot, 0, "I", 0)]
        ins = [int(v) for v in fb.np_vector(ot, 1, np.int32)]
        outs = [int(v) for v in fb.np_vector(ot, 2, np.int32)]
        operators.append(Operator(i, op, ins, outs, _options(fb, ot, op)))
    sigs: List[Signature] = []
    for st in fb.tables(model, 7):
        def tmap(idx):
            return {fb.string(m, 0): fb.scalar(m, 1, "I", 0) for m in fb.tables(st, idx)}
        sigs.append(Signature(fb.string(st, 2), tmap(0), tmap(1)))
    return Graph(tensors, operators, [int(v) for v in fb.np_vector(sub, 1, np.int32)],
                 [int(v) for v in fb.np_vector(sub, 2, np.int32)], sigs, fb.string(model, 3))
