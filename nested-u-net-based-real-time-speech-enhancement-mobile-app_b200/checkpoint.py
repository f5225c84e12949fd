"""Stream checkpoints: the history of running streams as one `.npz`, keyed by the reference's tensor names.

SURVEY 8(f)2: a stream must be able to leave the engine and continue elsewhere -- on another engine (another GPU, a
restarted process) or on a real TFLite signature runner.  A checkpoint therefore stores, per stream,

* every tensor of the signature under its OUTPUT name (`msfe6_ee_cur1`, `msfe6_en_h`, `state_c`, ... -- the dict the
  reference loop carries from call to call, `interpreter_proposed.py:215-350`, shapes `converter_proposed.py:729-867`;
  `nutls`: `interpreter_nunet_tls.py:36-289`),
* the frame loop's `in_buffer` / `out_buffer` (`interpreter_proposed.py:30-31, 203-204, 361-365`), so that wav-level
  streaming resumes sample-exactly,
* the attention rings when the engine was built with `stream_ctfa_history` (an extension; the reference graph has none).

Arrays are stacked over the saved streams: `msfe6_ee_cur1` is `[n_streams, 1, 1, 256, 64]`.
"""
from __future__ import annotations

import json
from typing import Dict, Iterable, List, Optional

import numpy as np

from ._lib import NUNET_VARIANT_DDB
from .interpreter import _engine_to_ref
from .state_table import STATE_SHAPES, STATE_SHAPES_DDB

FORMAT = "nunet_b200.stream_checkpoint/1"
FRAMING = ("in_buffer", "out_buffer")


def _shapes(variant_ddb: bool) -> Dict[str, tuple]:
    return STATE_SHAPES_DDB if variant_ddb else STATE_SHAPES


def _ring_names(engine) -> List[str]:
    names = []
    if getattr(engine, "stream_ctfa_history", False):
        i = 0
        while True:
            try:
                engine.state_numel(f"ctfa_ring{i}")
            except Exception:
                break
            names.append(f"ctfa_ring{i}")
            i += 1
    return names


def capture(engine, stream_ids: Optional[Iterable[int]] = None) -> Dict[str, np.ndarray]:
    """History of the given streams (default: all `max_streams`) as a dict of stacked arrays + a `meta` JSON string."""
    ids = list(range(engine.max_streams)) if stream_ids is None else [int(s) for s in stream_ids]
    ddb = engine.variant == NUNET_VARIANT_DDB
    shapes = _shapes(ddb)
    out: Dict[str, np.ndarray] = {}
    for name in engine.state_names():
        ref = _engine_to_ref(name, "cur")
        out[ref] = np.stack([engine.state_export(s, name).reshape(shapes[ref]) for s in ids])
    rings = _ring_names(engine)
    for name in list(FRAMING) + rings:
        out[name] = np.stack([engine.state_export(s, name) for s in ids])
    meta = {"format": FORMAT, "variant": "ddb" if ddb else "lstm", "streams": ids, "ctfa_mode": engine.ctfa_mode,
            "dc_mode": engine.dc_mode, "ctfa_rings": len(rings)}
    out["meta"] = np.array(json.dumps(meta))
    return out


def save(engine, path: str, stream_ids: Optional[Iterable[int]] = None) -> None:
    np.savez(path, **capture(engine, stream_ids))


def _meta(ck) -> dict:
    meta = json.loads(str(ck["meta"]))
    if meta.get("format") != FORMAT:
        raise ValueError(f"not a stream checkpoint (format {meta.get('format')!r})")
    return meta


def restore(engine, ck, stream_ids: Optional[Iterable[int]] = None) -> List[int]:
    """Write a captured/loaded checkpoint into `engine`; saved stream k goes to `stream_ids[k]` (default: the ids it was
    saved from).  Every tensor is validated against the engine's plan before anything is written."""
    meta = _meta(ck)
    ddb = engine.variant == NUNET_VARIANT_DDB
    if meta["variant"] != ("ddb" if ddb else "lstm"):
        raise ValueError(f"checkpoint holds {meta['variant']} streams, the engine runs {'ddb' if ddb else 'lstm'}")
    ids = list(meta["streams"]) if stream_ids is None else [int(s) for s in stream_ids]
    if len(ids) != len(meta["streams"]):
        raise ValueError(f"checkpoint holds {len(meta['streams'])} streams, {len(ids)} targets given")
    if any(s < 0 or s >= engine.max_streams for s in ids):
        raise ValueError("target stream id out of range")
    rings = _ring_names(engine)
    if len(rings) != meta["ctfa_rings"]:
        raise ValueError("checkpoint and engine disagree on stream_ctfa_history")
    plan = [(n, _engine_to_ref(n, "cur")) for n in engine.state_names()] + [(n, n) for n in list(FRAMING) + rings]
    for name, ref in plan:
        if ref not in ck:
            raise ValueError(f"checkpoint lacks tensor {ref}")
        a = np.asarray(ck[ref])
        if a.shape[0] != len(ids) or a[0].size != engine.state_numel(name):
            raise ValueError(f"{ref}: checkpoint shape {a.shape} does not fit the engine ({engine.state_numel(name)} values)")
    for name, ref in plan:
        a = np.asarray(ck[ref], dtype=np.float32)
        for k, s in enumerate(ids):
            engine.state_import(s, name, a[k])
    return ids


def load(engine, path: str, stream_ids: Optional[Iterable[int]] = None) -> List[int]:
    with np.load(path, allow_pickle=False) as ck:
        return restore(engine, {k: ck[k] for k in ck.files}, stream_ids)


def signature_feed(ck, k: int = 0) -> Dict[str, np.ndarray]:
    """The `*_prevK` / `_h` / `_c` keyword tensors a TFLite signature runner takes for saved stream `k` (everything
    but `input`): `runner(input=mag, **signature_feed(ck))` continues the stream on the reference interpreter."""
    meta = _meta(ck)
    shapes = _shapes(meta["variant"] == "ddb")
    feed = {}
    for ref, shape in shapes.items():
        feed[ref.replace("_cur", "_prev")] = np.asarray(ck[ref][k], dtype=np.float32).reshape(shape)
    return feed
