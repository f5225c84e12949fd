import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nunet_b200._lib import NUNET_VARIANT_LSTM_HYBRID
from nunet_b200.engine import NunetEngine
from nunet_b200.tflite_export import export_lstm_tflite, hybrid_weight_set
from nunet_b200.weights import load_default_weights, pack_blob
from oracle.tflite_graph import TFLiteGraph
w = load_default_weights()
export_lstm_tflite(w, "/tmp/dep.tflite")
hyb = TFLiteGraph("/tmp/dep.tflite", hybrid=True)
flt = TFLiteGraph("/tmp/dep.tflite", hybrid=False)
rng = np.random.default_rng(3)
x = (np.abs(rng.standard_normal((1, 1, 256, 1))) * 6).astype(np.float32)
feed = {k: torch.zeros(s) for k, s in hyb.input_shapes().items()}
feed["input"] = torch.from_numpy(x)
ref = hyb.run(feed)
rf = flt.run(feed)
eng = NunetEngine(pack_blob(hybrid_weight_set(w), 2), max_streams=1, variant=NUNET_VARIANT_LSTM_HYBRID)
eng.stream_reset()
y = eng.stream_step_mag(torch.from_numpy(x.reshape(1, 256)).cuda()).cpu().numpy()
print("model_out: engine vs hybrid", np.abs(y.reshape(-1) - ref["model_out"].numpy().reshape(-1)).max(), " hybrid vs float", np.abs(ref["model_out"] - rf["model_out"]).max().item())
for n in eng.state_names():
    import re
    m = re.fullmatch(r"(.+)_(\d+)", n)
    key = f"{m.group(1)}_cur{m.group(2)}" if m and not n.endswith(("_h", "_c")) else n
    a = eng.state_export(0, n)
    b = ref[key].numpy().reshape(-1)
    c = rf[key].numpy().reshape(-1)
    print(f"{key:22s} n={a.size:6d} |eng-hyb| {np.abs(a - b).max():9.3e}   |hyb-float| {np.abs(b - c).max():9.3e}   peak {np.abs(b).max():.3f}")
