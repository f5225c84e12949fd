"""A few streaming steps (S streams, plain launches) for ncu captures of the streaming kernels.  Not a benchmark."""
import os, sys
import numpy as np, torch
os.environ["NUNET_DEBUG_KNOBS"] = "1"
os.environ["NUNET_STREAM_GRAPH"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nunet_b200.engine import NunetEngine
from nunet_b200.synth import synth_clips
from nunet_b200.weights import load_default_weights, pack_blob
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
eng = NunetEngine(pack_blob(load_default_weights()), max_streams=S, ctfa_mode="frame_div32")
hops = torch.from_numpy(np.tile(synth_clips(32, 256 * 4), (S // 32 + 1, 1))[:S]).cuda()
out = torch.empty((S, 256), device="cuda")
eng.stream_reset()
for t in range(3):
    eng.stream_step_wav(hops[:, 256 * t:256 * (t + 1)].contiguous(), out)
torch.cuda.synchronize()
print("done", eng.last_launch_count)
