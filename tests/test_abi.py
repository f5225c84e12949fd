"""The C-ABI shared library: loads, exports exactly what include/nunet_b200.h declares, and fails loudly
(no CPU fallback) when no B200 is present.  No compute calls here."""
import ctypes
import os
import re

import numpy as np

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "nunet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nunet_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree():
    from nunet_b200 import _lib
    assert _declared() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol():
    from nunet_b200 import _lib
    L = _lib.lib()
    for name in _declared():
        assert hasattr(L, name), name
    assert L.nunet_abi_version() == 1
    assert [L.nunet_num_frames(n) for n in (0, 511, 512, 767, 768, 64000, 48000)] == [0, 0, 1, 1, 2, 249, 186]


def test_config_struct_layout():
    from nunet_b200._lib import NunetConfig
    assert ctypes.sizeof(NunetConfig) == 32


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from nunet_b200._lib import NunetError
    from nunet_b200.engine import NunetEngine
    with pytest.raises(NunetError) as ei:
        NunetEngine(b"NUNETW01" + b"\0" * 8, max_frames=4)
    assert ei.value.code == -4 and "CUDA" in str(ei.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "nunet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle/", "").lower() or f == "weights.py" or "import oracle" not in src, f
                assert "from oracle" not in src and "import oracle" not in src, f


def test_blob_is_parsed_and_packed_by_the_library(weights):
    """Host half of nunet_create (blob table, shape checks, weight packing incl. the up-sampling/inconv
    composition) runs without a GPU through nunet_blob_validate."""
    import ctypes as C
    from nunet_b200 import _lib
    from nunet_b200.weights import pack_blob
    L = _lib.lib()
    blob = pack_blob(weights)
    n = L.nunet_blob_validate(blob, len(blob), 0)
    assert n > 2_500_000, (n, L.nunet_last_error())
    assert L.nunet_blob_validate(blob[:1000], 1000, 0) == -1 and b"truncated" in L.nunet_last_error()
    bad = dict(weights)
    bad.pop("msfe4_de2_ta/kernel0")
    b2 = pack_blob(bad)
    assert L.nunet_blob_validate(b2, len(b2), 0) == -1 and b"msfe4_de2_ta/kernel0" in L.nunet_last_error()
    assert L.nunet_blob_validate(blob, len(blob), 1) == -1          # DDB variant not built yet: says so


def test_pybind11_layer_forwards_to_the_c_abi():
    """The thin pybind11 module (csrc/pybind_module.cpp) is the binding the package uses; on a CPU host its host-only entry
    points work and nunet_create fails loudly (no device) instead of falling back to anything."""
    from nunet_b200 import _lib
    from nunet_b200.weights import pack_blob, random_lstm_weights
    m = _lib.pyb()
    assert m.abi_version() == m.ABI_VERSION == _lib.lib().nunet_abi_version()
    assert [m.num_frames(n) for n in (511, 512, 64000)] == [0, 1, 249]
    blob = pack_blob(random_lstm_weights(0))
    assert m.blob_validate(blob, 0) > 2_500_000
    assert m.blob_validate(blob[:1000], 0) == -1 and "truncated" in m.last_error()
    import torch
    if not torch.cuda.is_available():
        rc, h = m.create(0, 0, 8, 0, 0, 1, 0, 0, blob)
        assert rc == -4 and h == 0 and "CUDA device" in m.last_error()
        from nunet_b200.engine import NunetEngine
        with pytest.raises(_lib.NunetError):
            NunetEngine(blob, max_frames=8)


def test_blob_parser_survives_corrupted_tables():
    """nunet_create / nunet_blob_validate take bytes from outside: a damaged table (count, rank, dimensions, offsets, names) must
    end in an error code with a message -- never in a crash, an out-of-bounds read or a wrap-around of the size arithmetic."""
    from nunet_b200 import _lib
    from nunet_b200.weights import pack_blob, random_lstm_weights
    m = _lib.pyb()
    blob = pack_blob(random_lstm_weights(0))
    cnt = int.from_bytes(blob[8:12], "little")
    assert m.blob_validate(blob, 0) > 0

    def patched(pos, value, width):
        b = bytearray(blob)
        b[pos:pos + width] = int(value).to_bytes(width, "little")
        return bytes(b)

    e0 = 16                                    # first table entry: name[64] | ndim u32 | dims u32[4] | offset u64 (floats)
    targeted = [
        (patched(8, 0xFFFFFFFF, 4), "truncated"),                      # count
        (patched(8, cnt + 1, 4), None),                                # table runs into the data
        (patched(e0 + 64, 5, 4), "rank"),                              # ndim
        (patched(e0 + 68, 0xFFFFFFFF, 4), "dimension"),                # one huge dimension
        (patched(e0 + 84, 2 ** 64 - 1, 8), "out of range"),            # offset + size would wrap
        (patched(e0 + 84, (len(blob) - 16 - cnt * 96) // 4, 8), "out of range"),   # starts at the very end
        (blob[:8] + b"\xff" * 8 + blob[16:], None),
        (b"NUNETW02" + blob[8:], "magic"),
        (blob[:16 + cnt * 96 - 1], "truncated"),
        (b"", "magic"),
    ]
    for b, word in targeted:
        assert m.blob_validate(b, 0) == -1
        assert word is None or word in m.last_error(), (word, m.last_error())
    # all four dimensions large: the element count must not wrap around 64 bits
    b = bytearray(blob)
    b[e0 + 64:e0 + 68] = (4).to_bytes(4, "little")
    for k in range(4):
        b[e0 + 68 + 4 * k:e0 + 72 + 4 * k] = (0x7FFFFFFF).to_bytes(4, "little")
    assert m.blob_validate(bytes(b), 0) == -1 and "out of range" in m.last_error()
    rng = np.random.default_rng(7)
    errors = 0
    for _ in range(60):
        b = bytearray(blob)
        base = 16 + int(rng.integers(cnt)) * 96
        field = int(rng.integers(4))
        if field == 0:
            b[base + int(rng.integers(64))] ^= 0xFF
        elif field == 1:
            b[base + 64:base + 68] = int(rng.integers(0, 2 ** 32)).to_bytes(4, "little")
        elif field == 2:
            k = int(rng.integers(4))
            b[base + 68 + 4 * k:base + 72 + 4 * k] = int(rng.integers(0, 2 ** 32)).to_bytes(4, "little")
        else:
            b[base + 84:base + 92] = int(rng.integers(0, 2 ** 63)).to_bytes(8, "little")
        r = m.blob_validate(bytes(b), 0)
        assert r == -1 or r > 0
        if r == -1:
            errors += 1
            assert len(m.last_error()) > 0
    assert errors > 20
