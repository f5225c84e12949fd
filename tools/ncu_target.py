"""One offline forward at a reduced batch, for ncu captures (`ncu -k regex:conv_tc ... python tools/ncu_target.py`).
Not a benchmark: numbers printed under a profiler are never bench values."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nunet_b200.engine import NunetEngine  # noqa: E402
from nunet_b200.synth import synth_clips  # noqa: E402
from nunet_b200.weights import load_default_weights, pack_blob  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = 64000
T = 1 + (N - 512) // 256
eng = NunetEngine(pack_blob(load_default_weights()), max_frames=B * T, device=0, ctfa_mode="causal_avg32")
wav = torch.from_numpy(synth_clips(8, N)).repeat((B + 7) // 8, 1)[:B].contiguous().cuda()
out = torch.empty((B, (T - 1) * 256 + 512), device="cuda")
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
for _ in range(reps):
    eng.forward_wav_into(wav, out)
torch.cuda.synchronize()
print("done", eng.last_launch_count)
