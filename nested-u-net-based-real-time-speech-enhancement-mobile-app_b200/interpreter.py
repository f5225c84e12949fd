"""Surface 2 of the reference, backed by the CUDA engine: the TFLite signature-runner contract

    it = Interpreter(model_path=...); it.allocate_tensors(); it.get_signature_list()
    run = it.get_signature_runner('nutls_lstm_sm')
    out = run(input=mag[1,1,256,1], msfe6_ee_prev1=..., ..., msfe6_de_c=...)   # dict with model_out + 130 tensors

(`dnn_model/interpreter_proposed.py:374-380, 215-350`; tensor names and shapes `converter_proposed.py:26-187`
inputs, `:729-867` outputs) and the frame loop `real_time_speech_enhancer` (`interpreter_proposed.py:15-370`).

The engine keeps the history resident on the GPU.  When the caller feeds back exactly the arrays the previous
call returned (what the reference loop does) nothing is imported; any other array is written into the engine
first, so foreign histories (e.g. from a real TFLite run) can be injected.
"""
from __future__ import annotations

import re
import time
from typing import Dict, List, Optional

import numpy as np
import torch

from .engine import NunetEngine
from .state_table import STATE_SHAPES
from .weights import expected_lstm_shapes, lstm_weights_from_h5, pack_blob, validate

SIGNATURE_KEYS = ("nutls_lstm_sm", "nutls_lstm")   # shipped file / converter script


def _engine_to_ref(name: str, which: str) -> str:
    """engine state name 'msfe4_ee2_3' -> 'msfe4_ee2_prev3' / 'msfe4_ee2_cur3'; LSTM names are unchanged."""
    m = re.fullmatch(r"(.+)_(\d+)", name)
    if m and not name.endswith(("_h", "_c")):
        return f"{m.group(1)}_{which}{m.group(2)}"
    return name


class SignatureRunner:
    def __init__(self, engine: NunetEngine):
        self._e = engine
        self._names: List[str] = engine.state_names()
        self._last_out: Dict[str, np.ndarray] = {}
        self._dev_in = torch.empty((1, 256), device=engine.device, dtype=torch.float32)
        for n in self._names:   # the engine's plan and the reference table must agree tensor by tensor
            ref = _engine_to_ref(n, "cur")
            if int(np.prod(STATE_SHAPES[ref])) != engine.state_numel(n):
                raise RuntimeError(f"history tensor {ref}: engine has {engine.state_numel(n)} values, "
                                   f"reference shape is {STATE_SHAPES[ref]}")

    def input_names(self) -> List[str]:
        return ["input"] + [_engine_to_ref(n, "prev") for n in self._names]

    def output_names(self) -> List[str]:
        return [_engine_to_ref(n, "cur") for n in self._names] + ["model_out"]

    def __call__(self, **kw) -> Dict[str, np.ndarray]:
        e = self._e
        if "input" not in kw:
            raise ValueError("missing signature input 'input'")
        expected = set(self.input_names())
        unknown = set(kw) - expected
        if unknown:
            raise ValueError(f"unknown signature inputs: {sorted(unknown)[:4]}")
        missing = expected - set(kw)
        if missing:
            raise ValueError(f"missing signature inputs: {sorted(missing)[:4]}")
        for n in self._names:
            given = kw[_engine_to_ref(n, "prev")]
            if given is self._last_out.get(_engine_to_ref(n, "cur")):
                continue                      # history already resident
            e.state_import(0, n, given)
        x = np.ascontiguousarray(kw["input"], dtype=np.float32).reshape(1, 256)
        self._dev_in.copy_(torch.from_numpy(x))
        y = e.stream_step_mag(self._dev_in)
        out: Dict[str, np.ndarray] = {}
        for n in self._names:
            ref = _engine_to_ref(n, "cur")
            out[ref] = e.state_export(0, n).reshape(STATE_SHAPES[ref])
        out["model_out"] = y.cpu().numpy().reshape(1, 1, 256, 1)
        self._last_out = out
        return out


class Interpreter:
    """`tf.lite.Interpreter` stand-in for the NUNet-TLS-LSTM graph.  `model_path` may be the reference `.h5`
    float checkpoint (or a role-named weight set via `weights=`)."""

    def __init__(self, model_path: Optional[str] = None, weights: Optional[dict] = None, device: int = 0,
                 num_threads: Optional[int] = None):
        if weights is None:
            if model_path is None:
                raise ValueError("model_path or weights required")
            if model_path.endswith(".tflite"):
                from .tflite_reader import lstm_weights_from_tflite
                weights = lstm_weights_from_tflite(model_path)
            else:
                weights = lstm_weights_from_h5(model_path)
        validate(weights, expected_lstm_shapes())
        self._blob = pack_blob(weights)
        self._device = device
        self._engine: Optional[NunetEngine] = None

    def allocate_tensors(self):
        if self._engine is None:
            self._engine = NunetEngine(self._blob, max_streams=1, device=self._device, dc_mode="edge")
            self._engine.stream_reset()

    def get_signature_list(self) -> dict:
        self.allocate_tensors()
        r = SignatureRunner(self._engine)
        return {SIGNATURE_KEYS[0]: {"inputs": r.input_names(), "outputs": r.output_names()}}

    def get_signature_runner(self, key: str = SIGNATURE_KEYS[0]) -> SignatureRunner:
        if key not in SIGNATURE_KEYS:
            raise ValueError(f"unknown signature key {key!r}; available: {SIGNATURE_KEYS[0]}")
        self.allocate_tensors()
        return SignatureRunner(self._engine)

    @property
    def engine(self) -> NunetEngine:
        self.allocate_tensors()
        return self._engine


def real_time_speech_enhancer(noisy_speech: np.ndarray, interpreter: Interpreter):
    """Frame loop of interpreter_proposed.py:15-370 with framing, network and overlap-add on the GPU:
    one `nunet_stream_step_wav_host` call per 256-sample hop.  Returns (enhanced, per-frame seconds)."""
    e = interpreter.engine
    e.stream_reset()
    frame_len, frame_step = 512, 256
    noisy = np.ascontiguousarray(noisy_speech, dtype=np.float32)
    num_blocks = (noisy.shape[0] - (frame_len - frame_step)) // frame_step
    out_file = np.zeros(len(noisy) + (frame_len - frame_step), np.float32)
    hop_out = np.empty((1, frame_step), np.float32)
    times = []
    for idx in range(num_blocks):
        t0 = time.time()
        e.stream_step_wav_host(noisy[None, idx * frame_step:(idx + 1) * frame_step], hop_out)
        out_file[idx * frame_step:(idx + 1) * frame_step] = hop_out[0]
        times.append(time.time() - t0)
    return out_file[frame_len - frame_step:], times
