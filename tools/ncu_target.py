"""One offline forward at a chosen batch, for ncu captures (`ncu -k regex:conv_tc3 ... python tools/ncu_target.py B`).
Not a benchmark: numbers printed under a profiler are never bench values.  Also writes the launch-name list of the
forward (gpurun_out/launch_names_b<B>.json) so that ncu's per-launch rows can be joined with the engine's unit names."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nunet_b200.engine import NunetEngine  # noqa: E402
from nunet_b200.synth import synth_clips  # noqa: E402
from nunet_b200.weights import load_default_weights, pack_blob  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = 64000
T = 1 + (N - 512) // 256
eng = NunetEngine(pack_blob(load_default_weights()), max_frames=B * T, device=0, ctfa_mode="causal_avg32")
wav = torch.from_numpy(synth_clips(8, N)).repeat((B + 7) // 8, 1)[:B].contiguous().cuda()
out = torch.empty((B, (T - 1) * 256 + 512), device="cuda")
eng.profile(True)
eng.forward_wav_into(wav, out)
torch.cuda.synchronize()
names = [(n, b) for n, _ms, b in eng.profile_entries()]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({"batch": B, "frames": B * T, "launches": names}, open(os.path.join(ROOT, "gpurun_out", f"launch_names_b{B}.json"), "w"))
print("done", eng.last_launch_count)
