"""Small offline + streaming run of both variants for compute-sanitizer (memcheck)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nunet_b200.engine import NunetEngine
from nunet_b200.synth import synth_clips
from nunet_b200.weights import pack_blob, random_ddb_weights, random_lstm_weights

for variant, w in ((0, random_lstm_weights(1)), (1, random_ddb_weights(1))):
    B, T = 3, 40    # long enough that most tiles lie inside one clip (tensor-map boxes, CTA pairs)
    eng = NunetEngine(pack_blob(w, variant), max_frames=B * T, max_streams=2, variant=variant)
    wav = torch.from_numpy(synth_clips(B, 512 + 256 * (T - 1))).cuda()
    y, est = eng.forward_wav(wav)
    eng.stream_reset()
    for t in range(6):    # 2 eager steps, 2 graph captures, 2 replays
        eng.stream_step_wav(torch.from_numpy(synth_clips(2, 256)).cuda())
    torch.cuda.synchronize()
    print("variant", variant, "ok", float(est.abs().max()))
    eng.close()
