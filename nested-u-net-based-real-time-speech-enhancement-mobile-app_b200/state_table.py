"""Names and shapes of the streaming history tensors of NUNet-TLS-LSTM, derived from the topology.

Reference tables: `dnn_model/interpreter_proposed.py:36-198` (zero dict, `*_curK` + LSTM states) and the
TensorSpecs of `converter_proposed.py:26-187`.  tests/test_oracle_pins.py checks this derivation against a
fixture extracted from those files (tests/golden/state_shapes_lstm.json).
"""
from __future__ import annotations

from typing import Dict, Tuple

from .weights import DEC_BLOCKS, ENC_BLOCKS

UNITS = 21


def state_prefixes(block: str) -> Tuple[str, str]:
    """'msfe4_en2' -> ('msfe4_ee2', 'msfe4_ed2'): conv-history / spconv-history prefixes."""
    head, tail = block.split("_")
    side, idx = tail[:2], tail[2:]
    a = "e" if side == "en" else "d"
    return f"{head}_{a}e{idx}", f"{head}_{a}d{idx}"


def _ddb(s: Dict[str, tuple], role: str, fb: int, c: int) -> None:
    """History of one dilated dense block (converter_nunet_tls.py:173-290 specs, :373-411 body): `in` / `out` keep one
    input row, layer k the last d = 2^(k-1) rows of its k*C/2-channel concatenated input."""
    h = c // 2
    s[f"{role}_cur_in"] = (1, 1, fb, c)
    for k in range(1, 7):
        s[f"{role}_cur{k}"] = (1, 1 << (k - 1), fb, k * h)
    s[f"{role}_cur_out"] = (1, 1, fb, h)


def _build(variant: str = "lstm") -> Dict[str, tuple]:
    s: Dict[str, tuple] = {}
    for side, blocks in (("en", ENC_BLOCKS), ("de", DEC_BLOCKS)):
        for block, f0, depth in blocks:
            pc, ps = state_prefixes(block)
            for k in range(1, depth + 1):
                f = f0 >> (k - 1)
                cin = (64 if k == 1 else 32) * (1 if side == "en" else 2)
                s[f"{pc}_cur{k}"] = (1, 1, f, cin)                            # input row of conv k
                s[f"{ps}_cur{k}"] = (1, 1, (f0 >> depth) << (k - 1), 64)      # input row of spconv k
            if variant == "ddb":
                _ddb(s, f"{block}_ddb", f0 >> depth, 32)
            else:
                s[f"{block}_h"] = s[f"{block}_c"] = (1, UNITS)
    if variant == "ddb":
        _ddb(s, "ddb", 4, 64)
    else:
        s["state_h"] = s["state_c"] = (1, UNITS)
    return s


STATE_SHAPES: Dict[str, tuple] = _build()
STATE_SHAPES_DDB: Dict[str, tuple] = _build("ddb")      # 208 tensors, 411 904 floats (interpreter_nunet_tls.py:36-289)
STATE_FLOATS = sum(int(a * b * c * d) if len(sh) == 4 else sh[1] for sh in STATE_SHAPES.values()
                   for (a, b, c, d) in [sh if len(sh) == 4 else (1, 1, 1, sh[1])])
