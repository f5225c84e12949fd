"""Evaluation caller of the path: `dnn_model/test_interface.py:51-110` with the test-set mapping of
`dataloader/dataloader.py:124-171` (`mapping_testset`) -- file-to-file enhancement and a per-SNR quality report.

Differences from the reference script, all forced by the offline image: PESQ / STOI (`pesq`, `pystoi`) are not
installable, so the report carries SI-SDR and segmental-free SNR gain instead (the `metric=` hook takes any
`f(clean, estimate, fs) -> float`, so PESQ/STOI plug in where they are available); wav files are read with a
dependency-free PCM16 reader (`soundfile` is absent).  The reference's per-bucket counters are reset inside its
loop (`test_interface.py:69`, a bug that makes every bucket average "sum / 1"); buckets are averaged properly here.

    python -m nunet_b200.evaluate --clean-dir data --noisy-dir data --weights log/saved_model/nutls_lstm.h5 [--out-dir enhanced]
"""
from __future__ import annotations

import argparse
import fnmatch
import os
import re
import struct
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np


# ----------------------------------------------------------------------------------------------- wav io (PCM16 mono)
def read_wav(path: str) -> Tuple[np.ndarray, int]:
    b = open(path, "rb").read()
    if b[:4] != b"RIFF" or b[8:12] != b"WAVE":
        raise ValueError(f"{path}: not a RIFF/WAVE file")
    pos, fmt, data = 12, None, None
    while pos + 8 <= len(b):
        cid, sz = b[pos:pos + 4], struct.unpack_from("<I", b, pos + 4)[0]
        if cid == b"fmt ":
            fmt = struct.unpack_from("<HHIIHH", b, pos + 8)
        elif cid == b"data":
            data = b[pos + 8:pos + 8 + sz]
        pos += 8 + sz + (sz & 1)
    if fmt is None or data is None or fmt[0] != 1 or fmt[5] != 16:
        raise ValueError(f"{path}: only PCM16 is supported")
    x = np.frombuffer(data, "<i2").astype(np.float64) / 32768.0
    if fmt[1] > 1:
        x = x.reshape(-1, fmt[1]).mean(axis=1)
    return x, int(fmt[2])


def write_wav(path: str, x: np.ndarray, fs: int) -> None:
    pcm = np.clip(np.round(np.asarray(x, np.float64) * 32768.0), -32768, 32767).astype("<i2").tobytes()
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(pcm)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, 1, fs, fs * 2, 2, 16))
        f.write(b"data" + struct.pack("<I", len(pcm)) + pcm)


# ----------------------------------------------------------------------------------------------- dataloader.py:11-15, 124-171
def min_max_norm(wav: np.ndarray, eps: float = 1e-8) -> np.ndarray:
    mx, mn = np.max(np.abs(wav)), np.min(np.abs(wav))
    return (wav - mn) / (mx - mn + eps)


class MappingTestset:
    """`mapping_testset`: pairs `<8-char id>_<snr>.wav` with `<id>.wav`, peak-normalises, clips to [-1, 1] and
    front-pads to a multiple of the hop."""

    def __init__(self, clean_dir: str, noisy_dir: str, hop_len: int = 256):
        self.clean_dir, self.noisy_dir, self.hop = clean_dir, noisy_dir, hop_len
        self.noisy_files = sorted(f for f in fnmatch.filter(os.listdir(noisy_dir), "*.wav") if re.search(r"_\d+\.wav$", f))

    def mapping_data(self):
        clean_list, noisy_list = [], []
        for nf in self.noisy_files:
            noisy, _ = read_wav(os.path.join(self.noisy_dir, nf))
            clean, _ = read_wav(os.path.join(self.clean_dir, nf[:8] + ".wav"))
            noisy, clean = np.clip(min_max_norm(noisy), -1, 1), np.clip(min_max_norm(clean), -1, 1)
            if len(noisy) % self.hop:
                pad = self.hop - len(noisy) % self.hop
                noisy, clean = np.pad(noisy, [pad, 0]), np.pad(clean, [pad, 0])
            clean_list.append(clean.astype(np.float32))
            noisy_list.append(noisy.astype(np.float32)[None])
        return clean_list, noisy_list

    def get_snr_index(self) -> List[str]:
        return ["".join(re.findall(r"\d+", f[8:])) for f in self.noisy_files]


def si_sdr(clean: np.ndarray, est: np.ndarray, fs: int = 16000) -> float:
    n = min(len(clean), len(est))
    ref, e = clean[:n] - clean[:n].mean(), est[:n] - est[:n].mean()
    a = np.dot(e, ref) / (np.dot(ref, ref) + 1e-12)
    tgt = a * ref
    return float(10 * np.log10(np.dot(tgt, tgt) / (np.dot(e - tgt, e - tgt) + 1e-12)))


def evaluate(model, clean_dir: str, noisy_dir: str, out_dir: Optional[str] = None, fs: int = 16000,
             metric: Callable[[np.ndarray, np.ndarray, int], float] = si_sdr) -> Dict[str, Dict[str, float]]:
    """`for noisy in noisy_list: pred = model(noisy, training=False)` (test_interface.py:57-63), scored per SNR bucket.
    Returns {snr: {"n", "input", "output"}} plus an "all" entry."""
    mt = MappingTestset(clean_dir, noisy_dir)
    clean_list, noisy_list = mt.mapping_data()
    snr_index = mt.get_snr_index()
    buckets: Dict[str, List[Tuple[float, float]]] = {}
    for i, noisy in enumerate(noisy_list):
        pred = np.asarray(model(noisy, training=False))[0]
        buckets.setdefault(snr_index[i], []).append((metric(clean_list[i], noisy[0], fs), metric(clean_list[i], pred, fs)))
        if out_dir:
            os.makedirs(out_dir, exist_ok=True)
            write_wav(os.path.join(out_dir, mt.noisy_files[i]), pred, fs)
    rep = {snr: {"n": len(v), "input": float(np.mean([a for a, _ in v])), "output": float(np.mean([b for _, b in v]))}
           for snr, v in sorted(buckets.items(), key=lambda kv: int(kv[0] or 0))}
    allv = [p for v in buckets.values() for p in v]
    rep["all"] = {"n": len(allv), "input": float(np.mean([a for a, _ in allv])), "output": float(np.mean([b for _, b in allv]))}
    return rep


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--clean-dir", required=True)
    ap.add_argument("--noisy-dir", required=True)
    ap.add_argument("--weights", required=True, help=".h5 (LSTM variant) or .tflite (either variant)")
    ap.add_argument("--variant", default="lstm", choices=["lstm", "ddb"])
    ap.add_argument("--out-dir", default=None)
    args = ap.parse_args()
    from . import models
    from .options import default_options
    cls = models.NUTLS if args.variant == "ddb" else models.NUTLS_LSTM
    model = cls(default_options()).build_model().load_weights(args.weights)
    rep = evaluate(model, args.clean_dir, args.noisy_dir, args.out_dir)
    print("###########################################")
    for snr, r in rep.items():
        print(f"# Testset performance [{snr}{'dB' if snr != 'all' else ''}]  files {r['n']}")
        print(f"# SI-SDR : {r['input']:.2f} dB -> {r['output']:.2f} dB")
    print("###########################################")


if __name__ == "__main__":
    main()
