"""Surface 1 of the reference, backed by the CUDA engine:

    m = models.NUTLS_LSTM(opt); model = m.build_model(); model.load_weights(path); y = model(x, training=False)

mirrors `dnn_model/models/proposed.py:13-22` (ctor), `:627-637` (build_model), `test_interface.py:45,58`
(load_weights / call) and `:639` (tflite_model: the one-frame model with zero history).  The returned objects
are plain callables, not Keras models; the arithmetic runs in csrc/*.cu (no CPU path).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from .engine import NunetEngine, num_frames
from ._lib import NUNET_VARIANT_DDB, NUNET_VARIANT_LSTM
from .weights import (VARIANT_DDB, VARIANT_LSTM, ddb_weights_from_tflite, expected_ddb_shapes, expected_lstm_shapes,
                      lstm_weights_from_h5, lstm_weights_from_tflite, pack_blob, validate)


class _Model:
    """What `build_model()` returns: waveform [B, N] -> enhanced waveform [B, (T-1)*256+512]."""

    def __init__(self, opt, ctfa_mode: str = "causal_avg32", variant: int = NUNET_VARIANT_LSTM):
        self.opt = opt
        self.ctfa_mode = ctfa_mode
        self.variant = variant
        self._weights: Optional[Dict[str, np.ndarray]] = None
        self._blob: Optional[bytes] = None
        self._engine: Optional[NunetEngine] = None
        self.device = int(getattr(opt, "device", 0))

    # Keras API subset ---------------------------------------------------------------------------
    def load_weights(self, path_or_set):
        """`.h5` path written by Keras save_weights (train_interface.py:99-100), a shipped `.tflite` (int8 tensors
        are dequantised; the only weight source of the dilated-dense variant), or a role-named weight set."""
        ddb = self.variant == NUNET_VARIANT_DDB
        if isinstance(path_or_set, str):
            if path_or_set.endswith(".tflite"):
                w = ddb_weights_from_tflite(path_or_set) if ddb else lstm_weights_from_tflite(path_or_set)
            elif ddb:
                raise ValueError("the dilated-dense variant ships no .h5 checkpoint: load nutls.tflite or a weight set")
            else:
                w = lstm_weights_from_h5(path_or_set)
        else:
            w = dict(path_or_set)
        validate(w, expected_ddb_shapes() if ddb else expected_lstm_shapes())
        self._weights, self._blob, self._engine = w, pack_blob(w, VARIANT_DDB if ddb else VARIANT_LSTM), None
        return self

    def save_weights(self, path: str):
        """Keras `model.save_weights(path)` (train_interface.py:99-100): the loaded weight set as a Keras-layout `.h5`
        (LSTM variant; `keras_export.export_lstm_h5`).  Loading the shipped `.tflite` and saving gives the dequantised
        checkpoint in the float file format."""
        if self._blob is None:
            raise RuntimeError("load_weights() first")
        if self.variant == NUNET_VARIANT_DDB:
            raise ValueError("the dilated-dense variant has no .h5 layout in the reference (its only checkpoint is nutls.tflite)")
        from .keras_export import export_lstm_h5
        export_lstm_h5(self._weights, path)

    def _get_engine(self, frames: int) -> NunetEngine:
        if self._blob is None:
            raise RuntimeError("load_weights() first (the engine has no random initialiser)")
        if self._engine is None or self._engine.max_frames < frames:
            if self._engine is not None:
                self._engine.close()
            self._engine = NunetEngine(self._blob, max_frames=frames, device=self.device, ctfa_mode=self.ctfa_mode,
                                       variant=self.variant)
        return self._engine

    def __call__(self, x, training: bool = False):
        if training:
            raise NotImplementedError("inference path only (training is out of scope, SURVEY 8)")
        as_numpy = not isinstance(x, torch.Tensor)
        xt = torch.as_tensor(np.asarray(x, dtype=np.float32)) if as_numpy else x.to(torch.float32)
        if xt.dim() != 2:
            raise ValueError("expected [batch, samples]")
        dev = torch.device("cuda", self.device)
        xt = xt.to(dev).contiguous()
        eng = self._get_engine(xt.shape[0] * num_frames(xt.shape[1]))
        y, _ = eng.forward_wav(xt, want_wav=True, want_mag=False)
        return y.cpu().numpy() if as_numpy else y

    predict = __call__

    def enhance_mag(self, x):
        """Estimated magnitude spectrogram [B, T, 257] (DC zero), the quantity the parity bar is stated on."""
        xt = torch.as_tensor(np.asarray(x, dtype=np.float32)).to(torch.device("cuda", self.device)).contiguous()
        eng = self._get_engine(xt.shape[0] * num_frames(xt.shape[1]))
        _, mag = eng.forward_wav(xt, want_wav=False, want_mag=True)
        return mag.cpu().numpy()


class _FrameModel:
    """What `tflite_model()` returns: [1,1,256,1] -> [1,1,256,1] with zero history (proposed.py:639-1151)."""

    def __init__(self, opt, variant: int = NUNET_VARIANT_LSTM):
        self.opt = opt
        self._blob = None
        self._weights = None
        self._engine = None
        self.variant = variant
        self.device = int(getattr(opt, "device", 0))

    def load_weights(self, path_or_set):
        ddb = self.variant == NUNET_VARIANT_DDB
        if isinstance(path_or_set, str):
            if path_or_set.endswith(".tflite"):
                w = ddb_weights_from_tflite(path_or_set) if ddb else lstm_weights_from_tflite(path_or_set)
            else:
                w = lstm_weights_from_h5(path_or_set)
        else:
            w = dict(path_or_set)
        validate(w, expected_ddb_shapes() if ddb else expected_lstm_shapes())
        self._weights = w
        self._blob = pack_blob(w, VARIANT_DDB if ddb else VARIANT_LSTM)
        self._engine = None          # created on first call: exporting needs no GPU
        return self

    def convert_to_tflite(self, path: str) -> dict:
        """The converter step on the other side of the path (converter_proposed.py:877-912: TFL_SIGNITURE around this
        per-frame model -> saved_model -> TFLiteConverter with Optimize.DEFAULT): writes the loaded weights as the reference's
        one-frame stateful graph with signature 'nutls_lstm_sm' and dynamic-range int8 weights (tflite_export.py).
        Host-only; no GPU needed."""
        if self._blob is None:
            raise RuntimeError("load_weights() first")
        if self.variant == NUNET_VARIANT_DDB:
            raise ValueError("only the NUNet-TLS-LSTM graph has an exporter (the dilated-dense variant ships no float checkpoint)")
        from .tflite_export import export_lstm_tflite
        return export_lstm_tflite(self._weights, path)

    def __call__(self, x, training: bool = False):
        x = np.asarray(x, dtype=np.float32).reshape(1, 256)
        if self._engine is None:
            self._engine = NunetEngine(self._blob, max_streams=1, device=self.device, variant=self.variant)
        self._engine.stream_reset()
        out = self._engine.stream_step_mag(torch.from_numpy(x).to(self._engine.device))
        return out.cpu().numpy().reshape(1, 1, 256, 1)


class NUTLS_LSTM:
    def __init__(self, opt):
        self.in_ch, self.mid_ch, self.out_ch = 1, 32, 64
        self.win_len, self.fft_len, self.hop_len = opt.win_len, opt.fft_len, opt.hop_len
        if (self.win_len, self.fft_len, self.hop_len) != (512, 512, 256):
            raise ValueError("the kernels are specialised for win_len = fft_len = 512, hop_len = 256")
        self.unit = 21
        self.opt = opt
        self.model = None

    def build_model(self) -> _Model:
        self.model = _Model(self.opt)
        return self.model

    def tflite_model(self) -> _FrameModel:
        return _FrameModel(self.opt)


class NUTLS:
    """The NUNet-TLS baseline with dilated-dense bottlenecks (`dnn_model/models/nunet_tls.py:13`, build_model :1007,
    tflite_model :1019)."""

    def __init__(self, opt):
        self.in_ch, self.mid_ch, self.out_ch = 1, 32, 64
        self.win_len, self.fft_len, self.hop_len = opt.win_len, opt.fft_len, opt.hop_len
        if (self.win_len, self.fft_len, self.hop_len) != (512, 512, 256):
            raise ValueError("the kernels are specialised for win_len = fft_len = 512, hop_len = 256")
        self.opt = opt
        self.model = None

    def build_model(self) -> _Model:
        self.model = _Model(self.opt, variant=NUNET_VARIANT_DDB)
        return self.model

    def tflite_model(self) -> _FrameModel:
        return _FrameModel(self.opt, variant=NUNET_VARIANT_DDB)
