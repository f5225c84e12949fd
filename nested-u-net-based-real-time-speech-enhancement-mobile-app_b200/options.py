"""`opt` namespace compatible with the reference's `options.Options().init(argparse...)` (options.py:10-73):
only the fields the model constructors read (models/proposed.py:18-22, :627-631) plus the artifact paths."""
from __future__ import annotations

import argparse


class Options:
    def init(self, parser: argparse.ArgumentParser) -> argparse.ArgumentParser:
        parser.add_argument("--batch_size", type=int, default=3)
        parser.add_argument("--test_batch", type=int, default=1)
        parser.add_argument("--arch", type=str, default="NUTLS-LSTM")
        parser.add_argument("--fft_len", type=int, default=512)
        parser.add_argument("--win_len", type=int, default=512)
        parser.add_argument("--hop_len", type=int, default=256)
        parser.add_argument("--fs", type=int, default=16000)
        parser.add_argument("--chunk_size", type=int, default=48000)
        parser.add_argument("--lstm_unit", type=int, default=21)
        parser.add_argument("--weight_path", type=str, default="./log/saved_model/nutls_lstm.h5")
        parser.add_argument("--tflite_path", type=str, default="./tflite/nutls_lstm.tflite")
        parser.add_argument("--device", type=int, default=0, help="CUDA device ordinal (B200)")
        return parser


def default_options(**overrides) -> argparse.Namespace:
    opt = Options().init(argparse.ArgumentParser()).parse_args([])
    for k, v in overrides.items():
        setattr(opt, k, v)
    return opt
