"""Generates nunet_b200/data/keras_layout_lstm.json from the reference's own weight file (run in the build container,
where /root/reference exists): the inventory a Keras `load_weights` walks -- `layer_names` in model order, `weight_names`
per layer, every dataset path with its shape and the role key of our weight set that fills it."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from nunet_b200.h5_reader import read_h5, read_h5_attrs  # noqa: E402
from nunet_b200.weights import lstm_weights_from_h5  # noqa: E402

REF = "/root/reference/dnn_model/log/saved_model/nutls_lstm.h5"
raw = read_h5(REF)
attrs = read_h5_attrs(REF)
trace = {}
roles = lstm_weights_from_h5(REF, trace=trace)
by_path = {p: k for k, p in trace.items()}
missing = [p for p in raw if p not in by_path]
assert not missing, missing[:5]
layout = {
    "source": "dnn_model/log/saved_model/nutls_lstm.h5 (Keras 2.12.0 save_weights)",
    "attrs": {g: a for g, a in attrs.items() if a},
    "groups": sorted(attrs),
    "datasets": [{"path": p, "shape": list(raw[p].shape), "role": by_path[p]} for p in raw],
}
out = os.path.join(ROOT, "nunet_b200", "data", "keras_layout_lstm.json")
json.dump(layout, open(out, "w"), indent=0)
print(len(layout["datasets"]), "datasets,", len(layout["groups"]), "groups ->", out, os.path.getsize(out), "bytes")
