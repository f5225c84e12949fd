"""Artifact emission: write a role-named NUNet-TLS-LSTM weight set back as the Keras weight file the
reference scripts load (`model.load_weights('./log/saved_model/nutls_lstm.h5')`, `test_interface.py:45`,
`converter_proposed.py:13`).

The file's inventory -- 340 `layer_names` in model order, `weight_names` per layer, 772 datasets under
`/<layer>/<layer>[/<cell>]/<var>:0` -- is not derivable from the role names alone (Keras numbers its anonymous
sub-layers globally: `layer_normalization_57`, `p_re_lu_12`, `lstm_cell_6`), so it is kept as a table extracted once from
the reference's own file (`data/keras_layout_lstm.json`, generator `tests/golden/make_keras_layout.py`)."""
from __future__ import annotations

import json
import os
from typing import Dict

import numpy as np

from .h5_writer import write_h5

_LAYOUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "keras_layout_lstm.json")


def load_layout() -> dict:
    with open(_LAYOUT) as f:
        return json.load(f)


def export_lstm_h5(weights: Dict[str, np.ndarray], path: str) -> None:
    """Write `weights` (role-named, as `weights.lstm_weights_from_h5` / `lstm_weights_from_tflite` return them) to `path`
    in the reference's `.h5` layout.  Raises if a tensor is missing or has the wrong number of elements."""
    layout = load_layout()
    datasets = {}
    for d in layout["datasets"]:
        if d["role"] not in weights:
            raise KeyError(f"weight set lacks {d['role']} (needed for {d['path']})")
        a = np.asarray(weights[d["role"]], dtype=np.float32)
        if a.size != int(np.prod(d["shape"])):
            raise ValueError(f"{d['role']}: {a.shape} does not fill {d['path']} {tuple(d['shape'])}")
        datasets[d["path"]] = a.reshape(d["shape"])
    attrs = {g: dict(layout["attrs"].get(g, {})) for g in layout["groups"]}
    write_h5(path, datasets, attrs)
