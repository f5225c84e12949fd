"""Python handle on the CUDA engine (csrc/engine.cu through the C ABI, bound by the thin pybind11 module
csrc/pybind_module.cpp).  torch is used only for device memory and streams; all arithmetic happens in the hand-written
sm_100a kernels."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import (NUNET_CTFA_CAUSAL_AVG32, NUNET_CTFA_FRAME_DIV32, NUNET_DC_EDGE, NUNET_DC_ZERO,
                   NUNET_VARIANT_LSTM, check)


def _p(t) -> int:
    """device address of a tensor, 0 for None"""
    return 0 if t is None else t.data_ptr()

CTFA_MODES = {"causal_avg32": NUNET_CTFA_CAUSAL_AVG32, "frame_div32": NUNET_CTFA_FRAME_DIV32}
DC_MODES = {"zero": NUNET_DC_ZERO, "edge": NUNET_DC_EDGE}


def num_frames(n_samples: int) -> int:
    """tf.signal.stft(frame_length=512, frame_step=256, pad_end=False) frame count (models/proposed.py:285)."""
    return 0 if n_samples < 512 else 1 + (n_samples - 512) // 256


def _host_ptr(a) -> Tuple[int, object]:
    """Pointer of a C-contiguous float32 host INPUT buffer (numpy array or CPU torch tensor; converted if needed)."""
    if isinstance(a, torch.Tensor):
        if a.is_cuda or a.dtype != torch.float32 or not a.is_contiguous():
            raise ValueError("host buffer must be a contiguous float32 CPU tensor")
        return a.data_ptr(), a
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a.ctypes.data, a


def _host_out(a, shape, what: str) -> int:
    """Pointer of a caller-supplied host OUTPUT buffer: it is written through a raw pointer, so it must already be
    float32, C-contiguous, writable and hold exactly `shape` elements (no silent copy, no short buffer)."""
    n = int(np.prod(shape))
    if isinstance(a, torch.Tensor):
        if a.is_cuda or a.dtype != torch.float32 or not a.is_contiguous() or a.numel() != n:
            raise ValueError(f"{what}: expected a contiguous float32 CPU tensor of {n} elements {tuple(shape)}")
        return a.data_ptr()
    if not (isinstance(a, np.ndarray) and a.dtype == np.float32 and a.flags.c_contiguous and a.flags.writeable
            and a.size == n):
        raise ValueError(f"{what}: expected a writable C-contiguous float32 array of {n} elements {tuple(shape)}")
    return a.ctypes.data


class NunetEngine:
    """One engine = one GPU, one packed weight set, fixed capacity (no allocation after construction)."""

    def __init__(self, blob: bytes, max_frames: int = 0, max_streams: int = 0, device: int = 0,
                 ctfa_mode: str = "causal_avg32", dc_mode: str = "edge", stream_ctfa_history: bool = False,
                 variant: int = NUNET_VARIANT_LSTM, chunk_frames: int = 0):
        self._L = _lib.pyb()
        self._h = 0
        self.device = torch.device("cuda", device)
        self.max_frames, self.max_streams = int(max_frames), int(max_streams)
        self.ctfa_mode, self.dc_mode = ctfa_mode, dc_mode
        self.variant, self.stream_ctfa_history = int(variant), bool(stream_ctfa_history)
        self.chunk_frames = int(chunk_frames)
        rc, h = self._L.create(int(variant), int(device), int(max_frames), int(max_streams), CTFA_MODES[ctfa_mode], DC_MODES[dc_mode],
                               int(bool(stream_ctfa_history)), int(chunk_frames), bytes(blob))
        check(rc)
        self._h = h

    def close(self):
        if getattr(self, "_h", 0):
            self._L.destroy(self._h)
            self._h = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _dev(self, t: torch.Tensor, shape=None, what: str = "tensor") -> torch.Tensor:
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
                and t.device == self.device):
            raise ValueError(f"{what}: expected a contiguous float32 tensor on {self.device}")
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise ValueError(f"{what}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
        return t

    @property
    def state_generation(self) -> int:
        """Changes whenever the history resident in the engine changes (step, reset, import)."""
        return check(self._L.state_generation(self._h))

    @property
    def last_launch_count(self) -> int:
        return self._L.last_launch_count(self._h)

    def profile(self, on: bool) -> None:
        check(self._L.profile_enable(self._h, int(on)))

    def profile_entries(self):
        """[(kernel name, device ms, algorithmic bytes)] of the last profiled forward/step call."""
        out = []
        for i in range(check(self._L.profile_count(self._h))):
            rc, name, ms, nb = self._L.profile_entry(self._h, i)
            check(rc)
            out.append((name, float(ms), float(nb)))
        return out

    # ------------------------------------------------------------------ offline
    def forward_wav(self, wav: torch.Tensor, want_wav: bool = True, want_mag: bool = True):
        """wav [B,N] (cuda) -> (enhanced wav [B,(T-1)*256+512] | None, est magnitudes [B,T,257] | None)."""
        wav = self._dev(wav, what="wav")
        if wav.dim() != 2:
            raise ValueError("wav must be [B, N]")
        B, N = wav.shape
        T = num_frames(N)
        if B < 1 or T < 1:
            raise _lib.NunetError(-1, "clip shorter than one 512-sample frame" if B >= 1 else "empty batch")
        out_wav = torch.empty((B, (T - 1) * 256 + 512), device=self.device, dtype=torch.float32) if want_wav else None
        out_mag = torch.empty((B, T, 257), device=self.device, dtype=torch.float32) if want_mag else None
        check(self._L.forward_wav_dev(self._h, wav.data_ptr(), B, N, _p(out_wav), _p(out_mag), self._stream()))
        return out_wav, out_mag

    def forward_wav_into(self, wav: torch.Tensor, out_wav: Optional[torch.Tensor], out_mag: Optional[torch.Tensor] = None):
        """Allocation-free variant for benchmarks."""
        wav = self._dev(wav, what="wav")
        if wav.dim() != 2:
            raise ValueError("wav must be [B, N]")
        B, N = wav.shape
        T = num_frames(N)
        if out_wav is not None:
            self._dev(out_wav, (B, (T - 1) * 256 + 512), "out_wav")
        if out_mag is not None:
            self._dev(out_mag, (B, T, 257), "out_mag")
        check(self._L.forward_wav_dev(self._h, wav.data_ptr(), B, N, _p(out_wav), _p(out_mag), self._stream()))

    def forward_mag(self, mag: torch.Tensor) -> torch.Tensor:
        """mag [B,T,256] (DC dropped) -> estimated magnitudes [B,T,256]."""
        mag = self._dev(mag, what="mag")
        if mag.dim() != 3 or mag.shape[2] != 256:
            raise ValueError("magnitudes must be [B, T, 256] (DC dropped)")
        B, T, F = mag.shape
        out = torch.empty_like(mag)
        check(self._L.forward_mag_dev(self._h, mag.data_ptr(), B, T, out.data_ptr(), self._stream()))
        return out

    def forward_wav_host(self, wav, out_wav=None, out_mag=None):
        """End-to-end host call: host wav [B,N] -> host buffers (H2D + kernels + D2H inside the call)."""
        p_in, keep = _host_ptr(wav)
        if keep.ndim != 2:
            raise ValueError("wav must be [B, N]")
        B, N = keep.shape
        T = num_frames(N)
        if B < 1 or T < 1:
            raise _lib.NunetError(-1, "clip shorter than one 512-sample frame" if B >= 1 else "empty batch")
        if out_wav is None:
            out_wav = np.empty((B, (T - 1) * 256 + 512), np.float32)
        p_out = _host_out(out_wav, (B, (T - 1) * 256 + 512), "out_wav")
        p_mag = _host_out(out_mag, (B, T, 257), "out_mag") if out_mag is not None else None
        check(self._L.forward_wav_host(self._h, p_in, B, N, p_out, p_mag or 0))
        return out_wav, out_mag

    def debug_read(self, name: str) -> np.ndarray:
        n = check(self._L.debug_read(self._h, name, 0, 0))
        buf = np.empty(n, np.float32)
        check(self._L.debug_read(self._h, name, buf.ctypes.data, n))
        return buf

    # ------------------------------------------------------------------ streaming
    def stream_reset(self, first: int = 0, count: Optional[int] = None):
        check(self._L.stream_reset(self._h, first, self.max_streams - first if count is None else count, self._stream()))

    def stream_step_mag(self, mag: torch.Tensor) -> torch.Tensor:
        mag = self._dev(mag, what="mag")
        if mag.dim() != 2 or mag.shape[1] != 256:
            raise ValueError(f"mag must be [S, 256], got {tuple(mag.shape)}")
        out = torch.empty_like(mag)
        check(self._L.stream_step_mag_dev(self._h, mag.data_ptr(), mag.shape[0], out.data_ptr(), self._stream()))
        return out

    def stream_step_wav(self, hop: torch.Tensor, out_hop: Optional[torch.Tensor] = None,
                        out_mag: Optional[torch.Tensor] = None) -> torch.Tensor:
        hop = self._dev(hop, what="hop")
        if hop.dim() != 2 or hop.shape[1] != 256:
            raise ValueError(f"hop must be [S, 256], got {tuple(hop.shape)}")
        if out_hop is None:
            out_hop = torch.empty_like(hop)
        else:
            self._dev(out_hop, hop.shape, "out_hop")
        if out_mag is not None:
            self._dev(out_mag, hop.shape, "out_mag")
        check(self._L.stream_step_wav_dev(self._h, hop.data_ptr(), hop.shape[0], out_hop.data_ptr(), _p(out_mag), self._stream()))
        return out_hop

    def stream_step_wav_host(self, hop, out_hop=None):
        p_in, keep = _host_ptr(hop)
        if keep.ndim != 2 or keep.shape[1] != 256:
            raise ValueError(f"hop must be [S, 256], got {tuple(keep.shape)}")
        if out_hop is None:
            out_hop = np.empty(keep.shape, np.float32)
        p_out = _host_out(out_hop, keep.shape, "out_hop")
        check(self._L.stream_step_wav_host(self._h, p_in, keep.shape[0], p_out))
        return out_hop

    # ------------------------------------------------------------------ history wire format
    def state_names(self) -> List[str]:
        out = []
        for i in range(check(self._L.state_count(self._h))):
            rc, name = self._L.state_name(self._h, i)
            check(rc)
            out.append(name)
        return out

    def state_numel(self, name: str) -> int:
        return check(self._L.state_numel(self._h, name))

    def state_export(self, stream_id: int, name: str) -> np.ndarray:
        buf = np.empty(self.state_numel(name), np.float32)
        check(self._L.state_export(self._h, stream_id, name, buf.ctypes.data))
        return buf

    def state_import(self, stream_id: int, name: str, value) -> None:
        a = np.ascontiguousarray(value, dtype=np.float32).reshape(-1)
        if a.size != self.state_numel(name):
            raise ValueError(f"{name}: expected {self.state_numel(name)} values, got {a.size}")
        check(self._L.state_import(self._h, stream_id, name, a.ctypes.data))

    def state_dict(self, stream_id: int = 0) -> Dict[str, np.ndarray]:
        return {n: self.state_export(stream_id, n) for n in self.state_names()}
