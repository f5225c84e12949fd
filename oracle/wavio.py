"""ORACLE-side helper (test infrastructure): minimal PCM16 mono WAV reader (soundfile is not installed)."""
import struct

import numpy as np


def read_wav(path):
    b = open(path, "rb").read()
    assert b[:4] == b"RIFF" and b[8:12] == b"WAVE"
    pos, fmt, data = 12, None, None
    while pos + 8 <= len(b):
        cid, sz = b[pos:pos + 4], struct.unpack_from("<I", b, pos + 4)[0]
        if cid == b"fmt ":
            fmt = struct.unpack_from("<HHIIHH", b, pos + 8)
        elif cid == b"data":
            data = b[pos + 8:pos + 8 + sz]
        pos += 8 + sz + (sz & 1)
    assert fmt is not None and fmt[0] == 1 and fmt[1] == 1 and fmt[5] == 16, fmt
    return np.frombuffer(data, "<i2").astype(np.float64) / 32768.0, fmt[2]
