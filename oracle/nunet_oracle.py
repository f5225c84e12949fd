"""ORACLE (test infrastructure, not product): CPU restatement of the reference NUNet-TLS-LSTM path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs
may import this file.  The product (`nunet_b200`) never does.

What it restates (all citations relative to the reference checkout, `dnn_model/`):
  * layer factories              models/proposed.py:125-282   (ctfa, conv, inconv, spconv, down/up_sampling)
  * offline graph                models/proposed.py:284-625   (`NUTLS_LSTM.train_model`)
  * one-frame stateful graph     converter_proposed.py:188-867 (`TFL_SIGNITURE.nutls_lstm`)
  * streaming frame loop         interpreter_proposed.py:15-370 (`real_time_speech_enhancer`)
The arithmetic itself lives in TensorFlow/Keras/TFLite (not vendored, version unpinned: README.md:81
says TF 2.9, the .h5 was written by Keras 2.12.0); Keras semantics are restated from their published
definitions (LayerNormalization non-fused path for eps < 1.001e-5, PReLU shared_axes, LSTM gate order
i,f,c,o, `tf.signal.stft/inverse_stft`, `padding='same'` rules).

PARITY PIN STATUS: the reference has no tests or golden outputs for this path and TensorFlow cannot
be imported here, so this oracle is pinned by (tests/test_oracle_pins.py): the literal window tables
of RTSE_NUTLS_LSTM.java:62-63, .h5 <-> .tflite weight agreement, the reference's state-shape tables,
the clean/noisy wav pairs (SI-SDR improvement), offline == streaming self-consistency, and the graph
oracle `oracle/tflite_graph.py`, which executes the reference's shipped `.tflite` flatbuffer itself.

Everything is written for clarity, not speed; tensors are NHWC like the reference.
"""
from __future__ import annotations

import math
import time
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

LN_EPS = 1e-8
UNITS = 21
TIME_SEQ = 32

ENC_BLOCKS = [("msfe6_en", 6), ("msfe5_en", 5), ("msfe4_en", 4), ("msfe4_en2", 4), ("msfe4_en3", 4), ("msfe3_en", 3)]
DEC_BLOCKS = [("msfe3_de", 3), ("msfe4_de", 4), ("msfe4_de2", 4), ("msfe4_de3", 4), ("msfe5_de", 5), ("msfe6_de", 6)]
DOWN_NAMES = ["msfe6_down_sampling", "msfe5_down_sampling", "msfe4_down_sampling",
              "msfe4_down_sampling2", "msfe4_down_sampling3", "msfe3_down_sampling"]
UP_NAMES = ["msfe3_upsampling", "msfe4_upsampling", "msfe4_upsampling2",
            "msfe4_upsampling3", "msfe5_upsampling", "msfe6_upsampling"]
# decoder block i pairs with encoder block 5-i (proposed.py:465,489,514,539,564,590)


def state_prefixes(block: str) -> Tuple[str, str, str]:
    """('msfe4_en2') -> conv-history prefix 'msfe4_ee2', spconv-history prefix 'msfe4_ed2', lstm prefix.

    Naming of converter_proposed.py:26-187."""
    head, tail = block.split("_")            # 'msfe4', 'en2'
    side, idx = tail[:2], tail[2:]
    a = "e" if side == "en" else "d"
    return f"{head}_{a}e{idx}", f"{head}_{a}d{idx}", block


class Oracle:
    """Float restatement of NUNet-TLS-LSTM.  `weights` is a role-named set (nunet_b200/weights.py)."""

    def __init__(self, weights: Dict[str, np.ndarray], dtype=torch.float32, ctfa_mode: str = "causal_avg32",
                 variant: str = "lstm"):
        """variant 'lstm' = NUNet-TLS-LSTM (models/proposed.py); 'ddb' = the NUNet-TLS baseline whose bottlenecks
        are dilated dense blocks (models/nunet_tls.py) -- everything else is shared."""
        assert ctfa_mode in ("causal_avg32", "frame_div32") and variant in ("lstm", "ddb")
        self.dt = dtype
        self.ctfa_mode = ctfa_mode
        self.variant = variant
        self.w = {k: torch.from_numpy(np.asarray(v, dtype=np.float32)).to(dtype) for k, v in weights.items()}

    # ------------------------------------------------------------------ Keras layer semantics
    def _ln_prelu(self, y: torch.Tensor, name: str) -> torch.Tensor:
        # LayerNormalization(epsilon=1e-8), axis=-1, non-fused path: tf.nn.moments + tf.nn.batch_normalization
        mean = y.mean(dim=-1, keepdim=True)
        var = ((y - mean) ** 2).mean(dim=-1, keepdim=True)
        inv = torch.rsqrt(var + LN_EPS) * self.w[f"{name}/gamma"]
        y = y * inv + (self.w[f"{name}/beta"] - mean * inv)
        # PReLU(shared_axes=[1,2,3]) -> one scalar alpha
        a = self.w[f"{name}/alpha"]
        return torch.where(y >= 0, y, a * y)

    def _conv2d(self, x: torch.Tensor, name: str, stride=(1, 1)) -> torch.Tensor:
        """Valid Conv2D on NHWC `x` with the Keras (kh,kw,Cin,Cout) kernel."""
        k = self.w[f"{name}/kernel"].permute(3, 2, 0, 1)
        y = F.conv2d(x.permute(0, 3, 1, 2), k, self.w[f"{name}/bias"], stride=stride)
        return y.permute(0, 2, 3, 1)

    @staticmethod
    def _with_history(x: torch.Tensor, prev: Optional[torch.Tensor]) -> torch.Tensor:
        """Offline: ZeroPadding2D(((1,0),..)) (proposed.py:200).  Streaming: Concatenate(axis=1)([prev, cur])
        in front of a `*_valid` layer (converter_proposed.py:226)."""
        if prev is None:
            return F.pad(x, (0, 0, 0, 0, 1, 0))
        return torch.cat([prev, x], dim=1)

    def inconv(self, x, name):  # proposed.py:218
        return self._ln_prelu(self._conv2d(x, name), name)

    def conv(self, x, name, prev=None):  # proposed.py:198 / conv_valid :208
        x = F.pad(self._with_history(x, prev), (0, 0, 1, 1))
        return self._ln_prelu(self._conv2d(x, name, stride=(1, 2)), name)

    def spconv(self, x, name, prev=None):  # proposed.py:227 / spconv_valid :240
        b, t, f, cin = x.shape
        y = self._conv2d(F.pad(self._with_history(x, prev), (0, 0, 1, 1)), name)   # [B,T,F,2*out_ch]
        out_ch = y.shape[-1] // 2
        y = y.reshape(b, -1, f, cin // 2, 2)       # Reshape((-1, F, Cin//2, 2)); -1 becomes 2T when out_ch=64
        y = y.permute(0, 1, 2, 4, 3)               # Permute((1,2,4,3))
        y = y.reshape(b, -1, f * 2, out_ch)        # Reshape((-1, 2F, out_ch))
        assert y.shape[1] == t
        return self._ln_prelu(y, name)

    def down_sampling(self, x, name):  # proposed.py:253  Conv2D (1,3) stride (1,2) 'same' -> pad right 1
        return self._conv2d(F.pad(x, (0, 0, 0, 1)), name, stride=(1, 2))

    def up_sampling(self, x, name):  # proposed.py:260  Conv2DTranspose (1,3) stride (1,2) 'same'
        k = self.w[f"{name}/kernel"].permute(3, 2, 0, 1)      # (kh,kw,Cout,Cin) -> (Cin,Cout,kh,kw)
        y = F.conv_transpose2d(x.permute(0, 3, 1, 2), k, self.w[f"{name}/bias"], stride=(1, 2))
        return y[..., : 2 * x.shape[2]].permute(0, 2, 3, 1)

    def _mlp(self, v, name):  # Conv1D/Conv2D(16,1) -> ReLU -> Conv(64,1) -> sigmoid
        h = torch.relu(v @ self.w[f"{name}/kernel0"] + self.w[f"{name}/bias0"])
        return torch.sigmoid(h @ self.w[f"{name}/kernel1"] + self.w[f"{name}/bias1"])

    def ctfa(self, x, block, ta_hist: Optional[torch.Tensor] = None):
        """proposed.py:125 (offline) / :162 `ctfa_rt` (one frame, no history => avg = TA/32).

        FA's input is TA broadcast over F, so FA is computed once per (b,t) (identical values).
        `ta_hist` [B,31,C] = TA of the 31 frames preceding this chunk (extension used only by the
        chunked self-consistency tests; the reference's offline graph starts from zeros)."""
        ta = self._mlp(x.mean(dim=2), f"{block}_ta")                     # [B,T,C]
        if self.ctfa_mode == "frame_div32":
            avg = ta / TIME_SEQ
            new_hist = None
        else:
            hist = ta_hist if ta_hist is not None else ta.new_zeros(ta.shape[0], TIME_SEQ - 1, ta.shape[2])
            padded = torch.cat([hist, ta], dim=1)                        # ZeroPadding2D((31,0)) at clip start
            avg = padded.unfold(1, TIME_SEQ, 1).mean(dim=-1)             # AveragePooling1D(32, strides=1)
            new_hist = padded[:, -(TIME_SEQ - 1):]
        fa = self._mlp(avg, f"{block}_fa")
        tfa = (fa * ta).unsqueeze(2)
        return x * tfa, new_hist

    def lstm_dense(self, x, lstm_name, dense_name, state=None):
        """Reshape [T, F*C] -> LSTM(21, return_sequences) -> Dense -> Reshape (proposed.py:305-309)."""
        b, t, f, c = x.shape
        seq = x.reshape(b, t, f * c)
        wk, wr, wb = (self.w[f"{lstm_name}/{n}"] for n in ("kernel", "recurrent_kernel", "bias"))
        h = seq.new_zeros(b, UNITS) if state is None else state[0]
        cst = seq.new_zeros(b, UNITS) if state is None else state[1]
        xw = seq @ wk + wb
        outs = []
        for i in range(t):
            z = xw[:, i] + h @ wr
            gi, gf, gc, go = z.split(UNITS, dim=1)                       # Keras order i, f, c, o
            cst = torch.sigmoid(gf) * cst + torch.sigmoid(gi) * torch.tanh(gc)
            h = torch.sigmoid(go) * torch.tanh(cst)
            outs.append(h)
        hs = torch.stack(outs, dim=1)
        y = hs @ self.w[f"{dense_name}/kernel"] + self.w[f"{dense_name}/bias"]
        return y.reshape(b, t, f, c), (h, cst)

    def ddb(self, x, role: str, st_in=None, st_out=None):
        """Dilated dense block (nunet_tls.py:190-272; use :383-410, main :678-700; one-frame form with explicit
        history converter_nunet_tls.py:373-411).  `in`: causal (2,3) conv C -> C/2 + PReLU; layers k = 1..6 with
        dilation d = 2^(k-1) in time AND frequency: grouped conv (groups = C/2, group g reads channels [g k, (g+1) k)
        of cat[out_{k-1}, .., out_0]) -> 1x1 conv -> LayerNorm -> PReLU; `out`: causal (2,3) conv C/2 -> C + PReLU.
        History: `in`/`out` keep their last input row, layer k the last d rows of its concatenated input."""
        def prev(sfx):
            return None if st_in is None else st_in[f"{role}_prev{sfx}"]

        def keep(sfx, full, rows):
            if st_out is not None:
                st_out[f"{role}_cur{sfx}"] = full[:, -rows:]

        def prelu(y, name):
            return torch.where(y >= 0, y, self.w[f"{name}/alpha"] * y)

        full = self._with_history(x, prev("_in"))
        keep("_in", full, 1)
        outs = [prelu(self._conv2d(F.pad(full, (0, 0, 1, 1)), f"{role}_in"), f"{role}_in")]
        for k in range(1, 7):
            d = 2 ** (k - 1)
            name = f"{role}_{k}"
            inp = torch.cat(outs[::-1], dim=3)                       # newest first: out_{k-1}, ..., out_0
            p = prev(str(k))
            full = F.pad(inp, (0, 0, 0, 0, d, 0)) if p is None else torch.cat([p, inp], dim=1)
            keep(str(k), full, d)
            z = F.pad(full, (0, 0, d, d)).permute(0, 3, 1, 2)        # ZeroPadding2D((d,0),(d,d)) / freq part
            kg = self.w[f"{name}/kernel0"].permute(3, 2, 0, 1)       # (2,3,k,h) -> (h, k, 2, 3)
            z = F.conv2d(z, kg, self.w[f"{name}/bias0"], dilation=(d, d), groups=kg.shape[0]).permute(0, 2, 3, 1)
            z = z @ self.w[f"{name}/kernel1"] + self.w[f"{name}/bias1"]
            outs.append(self._ln_prelu(z, name))
        full = self._with_history(outs[6], prev("_out"))
        keep("_out", full, 1)
        return prelu(self._conv2d(F.pad(full, (0, 0, 1, 1)), f"{role}_out"), f"{role}_out")

    # ------------------------------------------------------------------ the network
    def _msfe(self, block: str, depth: int, en_in, skips2, st_in, st_out, ctfa_hist, taps=None):
        """One nested sub-U-Net.  `skips2` = the paired encoder block's spconv outputs
        [de_1 .. de_n] (None on the encoder side).  Returns (ctfa_out + en_in, [de_1..de_n])."""
        pc, ps, pl = state_prefixes(block)

        def prev(prefix, k):
            return None if st_in is None else st_in[f"{prefix}_prev{k}"]

        def keep(prefix, k, v):
            if st_out is not None:
                st_out[f"{prefix}_cur{k}"] = v[:, -1:]

        ens = []
        cur = en_in
        for k in range(1, depth + 1):
            xin = cur if skips2 is None else torch.cat([cur, skips2[k - 1]], dim=3)
            keep(pc, k, xin)
            cur = self.conv(xin, f"{block}_conv{k}", prev(pc, k))
            ens.append(cur)
            if taps is not None:
                taps[f"{block}_conv{k}"] = cur
        if self.variant == "ddb":
            bb = self.ddb(cur, f"{block}_ddb", st_in, st_out)
        else:
            lstm_state = None if st_in is None else (st_in[f"{pl}_h"], st_in[f"{pl}_c"])
            bb, (h, c) = self.lstm_dense(cur, f"{block}_lstm", f"{block}_dense", lstm_state)
            if st_out is not None:
                st_out[f"{pl}_h"], st_out[f"{pl}_c"] = h, c
        if taps is not None:
            taps[f"{block}_bb"] = bb
        des = []
        cur = bb
        for k in range(1, depth + 1):
            xin = torch.cat([cur, ens[depth - k]], dim=3)
            keep(ps, k, xin)
            cur = self.spconv(xin, f"{block}_spconv{k}", prev(ps, k))
            des.append(cur)
            if taps is not None:
                taps[f"{block}_spconv{k}"] = cur
        gated, new_hist = self.ctfa(cur, block, None if ctfa_hist is None else ctfa_hist.get(block))
        if ctfa_hist is not None:
            ctfa_hist[block] = new_hist
        if taps is not None:
            taps[f"{block}_in"] = en_in
            taps[f"{block}_out"] = gated + en_in
        return gated + en_in, des[::-1]

    def net(self, mag, st_in=None, st_out=None, ctfa_hist=None, taps: Optional[dict] = None):
        """Magnitudes [B,T,256,1] (DC dropped) -> estimated magnitudes [B,T,256,1].

        st_in/st_out: reference-named history dicts (converter_proposed.py signature) or None for
        the zero-history offline graph."""
        x = self.inconv(mag, "input_layer")
        if taps is not None:
            taps["input_layer"] = x
        enc_de: List[List[torch.Tensor]] = []
        enc_out: List[torch.Tensor] = []
        for (block, depth), dn in zip(ENC_BLOCKS, DOWN_NAMES):
            en_in = self.inconv(x, f"{block}_in")
            out, des = self._msfe(block, depth, en_in, None, st_in, st_out, ctfa_hist, taps)
            x = self.down_sampling(out, dn)
            enc_de.append(des)
            enc_out.append(x)
            if taps is not None:
                taps[dn] = x
        if self.variant == "ddb":
            y = self.ddb(x, "ddb", st_in, st_out)
        else:
            main_state = None if st_in is None else (st_in["state_h"], st_in["state_c"])
            y, (h, c) = self.lstm_dense(x, "lstm", "dense", main_state)
            if st_out is not None:
                st_out["state_h"], st_out["state_c"] = h, c
        if taps is not None:
            taps["bb_main"] = y
        for i, ((block, depth), un) in enumerate(zip(DEC_BLOCKS, UP_NAMES)):
            j = 5 - i
            u = self.up_sampling(torch.cat([y, enc_out[j]], dim=3), un)
            en_in = self.inconv(u, f"{block}_in")
            y, _ = self._msfe(block, depth, en_in, enc_de[j], st_in, st_out, ctfa_hist, taps)
        return self._conv2d(y, "out_conv")

    # ------------------------------------------------------------------ offline surface (proposed.py:284-625)
    def stft(self, wav: torch.Tensor):
        """tf.signal.stft(frame_length=512, frame_step=256, fft_length=512, hann (periodic), pad_end=False)."""
        win = hann_window_periodic(512, wav.dtype)
        frames = wav.unfold(-1, 512, 256) * win
        spec = torch.fft.rfft(frames, n=512)
        return spec.abs(), torch.angle(spec)

    def forward_mag(self, mag257: torch.Tensor) -> torch.Tensor:
        """[B,T,257] magnitudes -> [B,T,257] estimated magnitudes with DC zero-padded (proposed.py:291,615-619)."""
        est = self.net(mag257[:, :, 1:, None].to(self.dt)).squeeze(3)
        return F.pad(est, (1, 0))

    def forward_wav(self, wav) -> Tuple[torch.Tensor, torch.Tensor]:
        """`model(noisy_wav)` of build_model (proposed.py:627): returns (wav_out [B,(T-1)*256+512], est_mag [B,T,257])."""
        wav = torch.as_tensor(wav, dtype=self.dt)
        mags, phase = self.stft(wav)
        est = self.forward_mag(mags)
        spec = torch.polar(est, phase)
        return inverse_stft(spec), est

    # ------------------------------------------------------------------ streaming surface
    def zero_state(self, batch: int = 1) -> Dict[str, torch.Tensor]:
        return {k: torch.zeros((batch,) + s[1:], dtype=self.dt) for k, s in state_shapes(self.variant).items()}

    def frame_step(self, inputs: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """The signature function `nutls_lstm` (converter_proposed.py:188-867): `input` + 104 `*_prevK`
        + 26 LSTM states -> `model_out` + 104 `*_curK` + 26 states."""
        out: Dict[str, torch.Tensor] = {}
        saved = self.ctfa_mode
        out["model_out"] = self.net(inputs["input"].to(self.dt), st_in=inputs, st_out=out)
        assert self.ctfa_mode == saved
        return out

    def real_time_speech_enhancer(self, noisy: np.ndarray, dc_pad: str = "edge", ctfa_hist=None,
                                  collect_mag: Optional[list] = None):
        """interpreter_proposed.py:15-370.  Returns (enhanced samples, per-frame seconds).

        numpy float64 FFTs like the reference (`np.fft.rfft` upcasts); model in `self.dt`."""
        frame_len, frame_step = 512, 256
        window = hann_window_periodic(frame_len, torch.float32).numpy().copy()
        window[0], window[-1] = 1e-7, 1e-7                                  # :22
        inv_window = inverse_stft_window(frame_len, frame_step, torch.float32).numpy()
        in_buffer = np.zeros(frame_len, np.float32)
        out_buffer = np.zeros(frame_len, np.float32)
        num_blocks = (noisy.shape[0] - (frame_len - frame_step)) // frame_step
        out_file = np.zeros(len(noisy) + (frame_len - frame_step))
        state = self.zero_state(1)
        times = []
        for idx in range(num_blocks):
            t0 = time.perf_counter()
            in_buffer[:-frame_step] = in_buffer[frame_step:]
            in_buffer[-frame_step:] = noisy[idx * frame_step:(idx + 1) * frame_step]
            spec = np.fft.rfft(in_buffer * window)
            in_mag, in_phase = np.abs(spec), np.angle(spec)
            sliced = in_mag.reshape(1, 1, -1, 1).astype(np.float32)[:, :, 1:]
            feed = {"input": torch.from_numpy(sliced)}
            feed.update({k.replace("_cur", "_prev"): v for k, v in state.items()})
            with torch.no_grad():
                res = self.frame_step_hist(feed, ctfa_hist) if ctfa_hist is not None else self.frame_step(feed)
            model_out = res.pop("model_out")
            state = res
            if collect_mag is not None:
                collect_mag.append(model_out.reshape(256).to(torch.float32).numpy().copy())
            est = np.pad(model_out.to(torch.float32).numpy(), ((0, 0), (0, 0), (1, 0), (0, 0)),
                         mode="edge" if dc_pad == "edge" else "constant").squeeze()
            block = np.fft.irfft(est * np.exp(1j * in_phase)).astype(np.float32) * inv_window
            out_buffer[:-frame_step] = out_buffer[frame_step:]
            out_buffer[-frame_step:] = 0
            out_buffer += block
            out_file[idx * frame_step:(idx + 1) * frame_step] = out_buffer[:frame_step]
            times.append(time.perf_counter() - t0)
        return out_file[frame_len - frame_step:], times

    def frame_step_hist(self, inputs, ctfa_hist: dict):
        """frame_step with carried CTFA history (extension: makes streaming == offline in causal_avg32 mode)."""
        out: Dict[str, torch.Tensor] = {}
        out["model_out"] = self.net(inputs["input"].to(self.dt), st_in=inputs, st_out=out, ctfa_hist=ctfa_hist)
        return out


# ---------------------------------------------------------------------- framing helpers
def hann_window_periodic(n: int, dtype=torch.float32) -> torch.Tensor:
    """tf.signal.hann_window(n) (periodic=True): 0.5 - 0.5 cos(2 pi k / n), evaluated in `dtype`."""
    k = torch.arange(n, dtype=dtype)
    return (0.5 - 0.5 * torch.cos(2.0 * math.pi * k / n)).to(dtype)


def inverse_stft_window(frame_len: int, hop: int, dtype=torch.float32) -> torch.Tensor:
    """tf.signal.inverse_stft_window_fn(hop, hann): w / sum over the overlapping shifts of w^2."""
    w = hann_window_periodic(frame_len, dtype)
    overlaps = -(-frame_len // hop)
    denom = F.pad(w * w, (0, overlaps * hop - frame_len)).reshape(overlaps, hop).sum(0, keepdim=True)
    denom = denom.repeat(overlaps, 1).reshape(-1)[:frame_len]
    return w / denom


def inverse_stft(spec: torch.Tensor) -> torch.Tensor:
    """tf.signal.inverse_stft(spec, 512, 256, 512, inverse_stft_window_fn(256)): irfft, window, overlap-add."""
    frames = torch.fft.irfft(spec, n=512) * inverse_stft_window(512, 256, spec.real.dtype)
    b, t, _ = frames.shape
    out = frames.new_zeros(b, (t - 1) * 256 + 512)
    for i in range(t):
        out[:, i * 256:i * 256 + 512] += frames[:, i]
    return out


def state_shapes(variant: str = "lstm") -> Dict[str, tuple]:
    """name -> shape of every `*_curK`/LSTM state (batch 1), derived from the topology; the test suite checks
    it against the tables in interpreter_proposed.py:36-198.  variant 'ddb': the 208 tensors of
    interpreter_nunet_tls.py:36-289 (LSTM states replaced by the dilated-dense histories)."""
    s: Dict[str, tuple] = {}

    def ddb(role, fb, c):
        s[f"{role}_cur_in"] = (1, 1, fb, c)
        for k in range(1, 7):
            s[f"{role}_cur{k}"] = (1, 1 << (k - 1), fb, k * c // 2)
        s[f"{role}_cur_out"] = (1, 1, fb, c // 2)

    f0s = {"msfe6": 256, "msfe5": 128, "msfe4_en": 64, "msfe4_en2": 32, "msfe4_en3": 16, "msfe3": 8,
           "msfe4_de": 16, "msfe4_de2": 32, "msfe4_de3": 64}
    for blocks, side in ((ENC_BLOCKS, "en"), (DEC_BLOCKS, "de")):
        for block, depth in blocks:
            head = block.split("_")[0]
            f0 = f0s.get(block, f0s.get(head))
            pc, ps, pl = state_prefixes(block)
            for k in range(1, depth + 1):
                f = f0 >> (k - 1)
                if side == "en":
                    cin = 64 if k == 1 else 32
                else:
                    cin = 128 if k == 1 else 64
                s[f"{pc}_cur{k}"] = (1, 1, f, cin)
                s[f"{ps}_cur{k}"] = (1, 1, (f0 >> depth) << (k - 1), 64)
            if variant == "ddb":
                ddb(f"{block}_ddb", f0 >> depth, 32)
            else:
                s[f"{pl}_h"] = s[f"{pl}_c"] = (1, UNITS)
    if variant == "ddb":
        ddb("ddb", 4, 64)
    else:
        s["state_h"] = s["state_c"] = (1, UNITS)
    return s


def min_max_norm(wav: np.ndarray, eps: float = 1e-8) -> np.ndarray:
    """dataloader.py:11-15 `minMaxNorm` followed by the clip of dataloader.py:71."""
    mx, mn = np.max(np.abs(wav)), np.min(np.abs(wav))
    return np.clip((wav - mn) / (mx - mn + eps), -1, 1)


def si_sdr(ref: np.ndarray, est: np.ndarray) -> float:
    n = min(len(ref), len(est))
    ref, est = ref[:n] - ref[:n].mean(), est[:n] - est[:n].mean()
    a = np.dot(est, ref) / (np.dot(ref, ref) + 1e-12)
    tgt = a * ref
    return float(10 * np.log10(np.dot(tgt, tgt) / (np.dot(est - tgt, est - tgt) + 1e-12)))
