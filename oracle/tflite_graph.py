"""ORACLE (test infrastructure, not product): float executor for the reference's SHIPPED TFLite graphs.

`dnn_model/tflite/nutls_lstm.tflite` and `nutls.tflite` are what the reference actually deploys
(`interpreter_proposed.py:374-380`, `RTSE_NUTLS_LSTM.java:571`).  No TFLite runtime can be installed here, so this
module executes the flatbuffers op by op (the 28 builtin operators they contain), driving them through their signature
exactly like `interpreter_proposed.py:215-350`: one frame in, 130 / 208 history tensors in and out.

It is the closest thing to "running the reference" that this container allows and pins the source restatement
(`oracle/nunet_oracle.py`) independently of our reading of the Keras code: identical (dequantised) weights must give
identical outputs to ~1e-5.  One deliberate difference from the real runtime: int8 weights are dequantised (q * scale)
and all arithmetic is float32, i.e. WITHOUT TFLite's dynamic-range quantisation of activations.

Operator semantics follow the TFLite reference kernels (tensorflow/lite/kernels/internal/reference); citations of
the graph structure are in SURVEY.md Appendix A.2.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

from nunet_b200.tflite_reader import Graph, Operator, read_tflite


def _same_pad(size: int, k: int, stride: int, dil: int):
    out = -(-size // stride)
    eff = (k - 1) * dil + 1
    total = max((out - 1) * stride + eff - size, 0)
    return total // 2, total - total // 2


class TFLiteGraph:
    """Executes one subgraph in float32.  `run(feed)` takes / returns tensors keyed by signature names."""

    def __init__(self, path: str, dtype=torch.float32, hybrid: bool = False):
        """hybrid=True emulates what the TFLite runtime actually does with these dynamic-range-quantised files
        (SURVEY 3A.4 #4, 8(f)3): CONV_2D / FULLY_CONNECTED ops whose weights are int8 run the HYBRID kernels -- the float
        input is quantised to int8 per call (per batch item for conv, per row for FC), the products are accumulated in
        int32 and rescaled by input scale x per-channel weight scale.  Restated from the published reference kernels
        (tensorflow/lite/kernels/conv.cc EvalHybridPerChannel, fully_connected.cc EvalHybrid,
        internal/reference/portable_tensor_utils.cc Asymmetric/SymmetricQuantizeFloats); no TFLite runtime exists here to
        pin it against, so it is a model of the deployed arithmetic, not a parity reference.  Default (False): float
        arithmetic on the dequantised weights, which is what the engine and O1 are tested against."""
        self.g: Graph = read_tflite(path)
        self.dt = dtype
        self.hybrid = hybrid
        self.sig = self.g.signatures[0]
        self.consts: Dict[int, torch.Tensor] = {}
        for t in self.g.tensors:
            if t.data is not None:
                if t.dtype in (np.int32, np.int64):
                    self.consts[t.index] = torch.from_numpy(t.data.astype(np.int64))
                elif t.dtype == np.float32:
                    self.consts[t.index] = torch.from_numpy(t.data).to(dtype)
                # int8 constants are materialised lazily through `weight()` (hybrid kernels) or DEQUANTIZE

    # ---------------------------------------------------------------- helpers
    def weight(self, idx: int) -> torch.Tensor:
        if idx not in self.consts:
            self.consts[idx] = torch.from_numpy(self.g.tensors[idx].dequantized()).to(self.dt)
        return self.consts[idx]

    # ---------------------------------------------------------------- hybrid (dynamic-range) kernels
    def _is_int8(self, idx: int) -> bool:
        t = self.g.tensors[idx]
        return self.hybrid and t.data is not None and t.dtype == np.int8

    def _wq(self, idx: int):
        """int8 weight values (as float64, exact) and per-output-channel scales (float32) of tensor idx."""
        t = self.g.tensors[idx]
        scale = np.asarray(t.scale, np.float32).reshape(-1)
        n_out = t.shape[0]
        if scale.size == 1:
            scale = np.full(n_out, scale[0], np.float32)
        return torch.from_numpy(t.data.astype(np.float64)), torch.from_numpy(scale)

    @staticmethod
    def _round_away(x: torch.Tensor) -> torch.Tensor:      # TfLiteRound = std::round: halves away from zero
        return torch.sign(x) * torch.floor(torch.abs(x) + 0.5)

    @classmethod
    def _quantize_rows(cls, x2: torch.Tensor, asymmetric: bool):
        """x2 [rows, n] float32 -> (q - zero_point as float64 [rows, n], scale float32 [rows]) per row, like
        Asymmetric/SymmetricQuantizeFloats (scale and zero point derived in double, values quantised in float)."""
        x2 = x2.to(torch.float32)
        if asymmetric:
            rmin = torch.clamp(x2.min(dim=1).values.double(), max=0.0)
            rmax = torch.clamp(x2.max(dim=1).values.double(), min=0.0)
            flat = rmin == rmax
            scale = torch.where(flat, torch.ones_like(rmin), (rmax - rmin) / 255.0)
            zp_min, zp_max = -128.0 - rmin / scale, 127.0 - rmax / scale
            err_min, err_max = 128.0 + (rmin / scale).abs(), 127.0 + (rmax / scale).abs()
            zp = torch.where(err_min < err_max, zp_min, zp_max)
            zp = torch.where(zp <= -128.0, torch.full_like(zp, -128.0),
                             torch.where(zp >= 127.0, torch.full_like(zp, 127.0), cls._round_away(zp)))
            zp = torch.where(flat, torch.zeros_like(zp), zp)
            scale_f = scale.to(torch.float32)
            inv = (1.0 / scale_f).to(torch.float32)
            q = cls._round_away(zp.to(torch.float32)[:, None] + x2 * inv[:, None]).clamp(-128, 127)
            q = torch.where(flat[:, None], torch.zeros_like(q), q)
            return q.double() - zp[:, None], scale_f
        rng = x2.abs().max(dim=1).values
        flat = rng == 0
        scale_f = torch.where(flat, torch.ones_like(rng), rng / 127.0)
        inv = torch.where(flat, torch.zeros_like(rng), 127.0 / rng)
        q = cls._round_away(x2 * inv[:, None]).clamp(-127, 127)
        return q.double(), scale_f

    def input_shapes(self) -> Dict[str, tuple]:
        return {k: self.g.tensors[i].shape for k, i in self.sig.inputs.items()}

    def output_shapes(self) -> Dict[str, tuple]:
        return {k: self.g.tensors[i].shape for k, i in self.sig.outputs.items()}

    @staticmethod
    def _act(y, act):
        if act == 0:
            return y
        if act == 1:
            return torch.relu(y)
        raise NotImplementedError(f"fused activation {act}")

    # ---------------------------------------------------------------- execution
    def run(self, feed: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        v: Dict[int, torch.Tensor] = {}
        for name, idx in self.sig.inputs.items():
            v[idx] = torch.as_tensor(feed[name]).to(self.dt)

        def get(i: int) -> Optional[torch.Tensor]:
            if i < 0:
                return None
            if i in v:
                return v[i]
            if i in self.consts:
                return self.consts[i]
            t = self.g.tensors[i]
            if t.data is not None:
                return self.weight(i)
            raise KeyError(f"tensor {i} ({t.name}) has no value")

        def ints(i: int) -> List[int]:
            return [int(x) for x in get(i).reshape(-1).tolist()]

        for op in self.g.operators:
            self._exec(op, v, get, ints)
        return {name: v[idx] for name, idx in self.sig.outputs.items()}

    def _exec(self, op: Operator, v, get, ints):
        k, o, a = op.op, op.outputs, op.inputs
        opt = op.options
        if k == "CONV_2D" and self._is_int8(a[1]):
            # hybrid per-channel conv: the input is quantised per batch item (asymmetric), padding is the zero point
            x, b = get(a[0]), get(a[2]) if len(a) > 2 else None
            wq, wscale = self._wq(a[1])
            sh, sw, dh, dw = opt["stride_h"], opt["stride_w"], opt["dilation_h"], opt["dilation_w"]
            groups = x.shape[3] // wq.shape[3]
            xq, xscale = self._quantize_rows(x.reshape(x.shape[0], -1), asymmetric=True)
            xi = xq.reshape(x.shape).permute(0, 3, 1, 2)
            if opt["padding"] == 0:   # SAME
                pt, pb = _same_pad(x.shape[1], wq.shape[1], sh, dh)
                pl, pr = _same_pad(x.shape[2], wq.shape[2], sw, dw)
                xi = F.pad(xi, (pl, pr, pt, pb))
            acc = F.conv2d(xi, wq.permute(0, 3, 1, 2), None, stride=(sh, sw), dilation=(dh, dw), groups=groups)   # exact in float64
            y = acc.to(torch.float32) * (xscale.reshape(-1, 1, 1, 1) * wscale.reshape(1, -1, 1, 1))
            if b is not None:
                y = y + b.to(torch.float32).reshape(1, -1, 1, 1)
            v[o[0]] = self._act(y.permute(0, 2, 3, 1).to(self.dt), opt["act"])
        elif k == "FULLY_CONNECTED" and self._is_int8(a[1]):
            x = get(a[0])
            b = get(a[2]) if len(a) > 2 and a[2] >= 0 else None
            wq, wscale = self._wq(a[1])
            xq, xscale = self._quantize_rows(x.reshape(-1, wq.shape[1]), asymmetric=bool(opt.get("asymmetric_quantize_inputs")))
            y = (xq @ wq.t()).to(torch.float32) * (xscale[:, None] * wscale[None, :])
            if b is not None:
                y = y + b.to(torch.float32)
            if opt.get("keep_num_dims"):
                y = y.reshape(*x.shape[:-1], wq.shape[0])
            v[o[0]] = self._act(y.to(self.dt), opt.get("act", 0))
        elif k == "CONV_2D":
            x, w, b = get(a[0]), self.weight(a[1]), get(a[2]) if len(a) > 2 else None
            sh, sw, dh, dw = opt["stride_h"], opt["stride_w"], opt["dilation_h"], opt["dilation_w"]
            groups = x.shape[3] // w.shape[3]
            xi = x.permute(0, 3, 1, 2)
            if opt["padding"] == 0:   # SAME
                pt, pb = _same_pad(x.shape[1], w.shape[1], sh, dh)
                pl, pr = _same_pad(x.shape[2], w.shape[2], sw, dw)
                xi = F.pad(xi, (pl, pr, pt, pb))
            y = F.conv2d(xi, w.permute(0, 3, 1, 2), b, stride=(sh, sw), dilation=(dh, dw), groups=groups)
            v[o[0]] = self._act(y.permute(0, 2, 3, 1), opt["act"])
        elif k == "TRANSPOSE_CONV":
            out_shape, w, x = ints(a[0]), get(a[1]), get(a[2])
            b = get(a[3]) if len(a) > 3 else None
            sh, sw = opt["stride_h"], opt["stride_w"]
            # weights [Cout, kh, kw, Cin] -> torch conv_transpose2d weight (Cin, Cout, kh, kw)
            y = F.conv_transpose2d(x.permute(0, 3, 1, 2), w.permute(3, 0, 1, 2), None, stride=(sh, sw))
            H, W = out_shape[1], out_shape[2]
            if opt["padding"] == 0:   # SAME: crop the full output symmetrically-left like the TFLite reference kernel
                pt = max((x.shape[1] - 1) * sh + w.shape[1] - H, 0) // 2
                pl = max((x.shape[2] - 1) * sw + w.shape[2] - W, 0) // 2
            else:
                pt = pl = 0
            y = y[:, :, pt:pt + H, pl:pl + W]
            if b is not None:
                y = y + b.reshape(1, -1, 1, 1)
            v[o[0]] = y.permute(0, 2, 3, 1)
        elif k == "FULLY_CONNECTED":
            x, w = get(a[0]), self.weight(a[1])
            b = get(a[2]) if len(a) > 2 and a[2] >= 0 else None
            y = x.reshape(-1, w.shape[1]) @ w.t()
            if b is not None:
                y = y + b
            if opt.get("keep_num_dims"):
                y = y.reshape(*x.shape[:-1], w.shape[0])
            v[o[0]] = self._act(y, opt.get("act", 0))
        elif k in ("ADD", "SUB", "MUL"):
            x, y = get(a[0]), get(a[1])
            if x.dtype != y.dtype:
                x, y = (x.to(self.dt), y.to(self.dt)) if self.dt in (x.dtype, y.dtype) else (x, y.to(x.dtype))
            r = x + y if k == "ADD" else (x - y if k == "SUB" else x * y)
            v[o[0]] = self._act(r, opt.get("act", 0))
        elif k == "SQUARED_DIFFERENCE":
            d = get(a[0]) - get(a[1])
            v[o[0]] = d * d
        elif k == "RSQRT":
            v[o[0]] = torch.rsqrt(get(a[0]))
        elif k == "LOGISTIC":
            v[o[0]] = torch.sigmoid(get(a[0]))
        elif k == "TANH":
            v[o[0]] = torch.tanh(get(a[0]))
        elif k == "PRELU":
            x, al = get(a[0]), get(a[1])
            v[o[0]] = torch.where(x >= 0, x, al * x)
        elif k == "MEAN":
            axes = [ax % get(a[0]).dim() for ax in ints(a[1])]
            v[o[0]] = get(a[0]).mean(dim=axes, keepdim=opt["keep_dims"])
        elif k == "REDUCE_PROD":
            x = get(a[0])
            for ax in sorted((ax % x.dim() for ax in ints(a[1])), reverse=True):
                x = x.prod(dim=ax, keepdim=opt["keep_dims"])
            v[o[0]] = x
        elif k == "AVERAGE_POOL_2D":
            x = get(a[0]).permute(0, 3, 1, 2)
            if opt["padding"] == 0:
                pt, pb = _same_pad(x.shape[2], opt["filter_h"], opt["stride_h"], 1)
                pl, pr = _same_pad(x.shape[3], opt["filter_w"], opt["stride_w"], 1)
                if pt or pb or pl or pr:
                    raise NotImplementedError("SAME average pool with real padding")
            y = F.avg_pool2d(x, (opt["filter_h"], opt["filter_w"]), (opt["stride_h"], opt["stride_w"]))
            v[o[0]] = self._act(y.permute(0, 2, 3, 1), opt["act"])
        elif k == "CONCATENATION":
            v[o[0]] = self._act(torch.cat([get(i) for i in a], dim=opt["axis"]), opt.get("act", 0))
        elif k == "RESHAPE":
            shape = ints(a[1]) if len(a) > 1 and a[1] >= 0 else [int(s) for s in opt["new_shape"]]
            v[o[0]] = get(a[0]).reshape(shape)
        elif k == "TRANSPOSE":
            v[o[0]] = get(a[0]).permute(ints(a[1])).contiguous()
        elif k == "PAD":
            x, pads = get(a[0]), get(a[1]).reshape(-1, 2).tolist()
            flat = []
            for before, after in reversed(pads):
                flat += [int(before), int(after)]
            v[o[0]] = F.pad(x, flat)
        elif k == "EXPAND_DIMS":
            v[o[0]] = get(a[0]).unsqueeze(ints(a[1])[0])
        elif k == "SHAPE":
            v[o[0]] = torch.tensor(list(get(a[0]).shape), dtype=torch.int64)
        elif k == "PACK":
            v[o[0]] = torch.stack([get(i) for i in a], dim=opt["axis"])
        elif k == "UNPACK":
            for j, t in enumerate(torch.unbind(get(a[0]), dim=opt["axis"])):
                v[o[j]] = t
        elif k == "SPLIT":
            axis, x = ints(a[0])[0], get(a[1])
            for j, t in enumerate(torch.chunk(x, opt["num_splits"], dim=axis)):
                v[o[j]] = t
        elif k == "GATHER":
            x, idx = get(a[0]), get(a[1])
            v[o[0]] = torch.index_select(x, opt["axis"], idx.reshape(-1).long()).reshape(
                *x.shape[:opt["axis"]], *idx.shape, *x.shape[opt["axis"] + 1:])
        elif k == "STRIDED_SLICE":
            x = get(a[0])
            begin, end, strides = ints(a[1]), ints(a[2]), ints(a[3])
            if opt["ellipsis_mask"] or opt["new_axis_mask"]:
                raise NotImplementedError("strided_slice ellipsis / new_axis")
            sl, squeeze = [], []
            for d in range(len(begin)):
                b_ = None if (opt["begin_mask"] >> d) & 1 else begin[d]
                e_ = None if (opt["end_mask"] >> d) & 1 else end[d]
                if (opt["shrink_axis_mask"] >> d) & 1:
                    bb = begin[d] if begin[d] >= 0 else begin[d] + x.shape[d]
                    sl.append(slice(bb, bb + 1, 1))
                    squeeze.append(d)
                else:
                    if strides[d] <= 0:
                        raise NotImplementedError("negative stride")
                    sl.append(slice(b_, e_, strides[d]))
            y = x[tuple(sl)]
            for d in reversed(squeeze):
                y = y.squeeze(d)
            v[o[0]] = y
        elif k == "SPACE_TO_BATCH_ND":
            x, block, pads = get(a[0]), ints(a[1]), get(a[2]).reshape(-1, 2).tolist()
            if len(block) != 2:
                raise NotImplementedError("space_to_batch rank")
            x = F.pad(x, (0, 0, int(pads[1][0]), int(pads[1][1]), int(pads[0][0]), int(pads[0][1])))
            n, h, w, c = x.shape
            bh, bw = block
            x = x.reshape(n, h // bh, bh, w // bw, bw, c).permute(2, 4, 0, 1, 3, 5)
            v[o[0]] = x.reshape(n * bh * bw, h // bh, w // bw, c)
        elif k == "BATCH_TO_SPACE_ND":
            x, block, crops = get(a[0]), ints(a[1]), get(a[2]).reshape(-1, 2).tolist()
            bh, bw = block
            nb, h, w, c = x.shape
            n = nb // (bh * bw)
            x = x.reshape(bh, bw, n, h, w, c).permute(2, 3, 0, 4, 1, 5).reshape(n, h * bh, w * bw, c)
            v[o[0]] = x[:, int(crops[0][0]):h * bh - int(crops[0][1]), int(crops[1][0]):w * bw - int(crops[1][1]), :]
        elif k == "DEQUANTIZE":
            v[o[0]] = self.weight(a[0])
        else:
            raise NotImplementedError(f"operator {k}")


def zero_feed(graph: TFLiteGraph) -> Dict[str, torch.Tensor]:
    return {k: torch.zeros(s, dtype=graph.dt) for k, s in graph.input_shapes().items()}


def stream_frames(graph: TFLiteGraph, mags: np.ndarray, input_name: str = "input", output_name: str = "model_out"):
    """Drive the signature frame by frame like interpreter_proposed.py:215-350: `*_curK` outputs feed the `*_prevK`
    inputs of the next call, LSTM `_h/_c` feed themselves.  mags [T,256] -> estimated magnitudes [T,256]."""
    feed = zero_feed(graph)
    outs = []
    with torch.no_grad():
        for t in range(mags.shape[0]):
            feed[input_name] = torch.from_numpy(mags[t].reshape(1, 1, 256, 1).astype(np.float32))
            res = graph.run(feed)
            outs.append(res[output_name].reshape(256).to(torch.float32).numpy().copy())
            for k, val in res.items():
                if k == output_name:
                    continue
                kin = k.replace("_cur", "_prev")
                if kin in feed:
                    feed[kin] = val
    return np.stack(outs)
