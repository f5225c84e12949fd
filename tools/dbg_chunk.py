import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nunet_b200.engine import NunetEngine
from nunet_b200.synth import synth_clips
from nunet_b200.weights import load_default_weights, pack_blob
blob = pack_blob(load_default_weights())
B, T = 3, 12
wav = torch.from_numpy(synth_clips(B, 512 + 256 * (T - 1), first_clip=500)).cuda()
for mode in ("frame_div32",):
    whole = NunetEngine(blob, max_frames=B * T, ctfa_mode=mode)
    y0, e0 = whole.forward_wav(wav)
    for chunk in (1, 2, 3):
        cut = NunetEngine(blob, max_frames=B * T, ctfa_mode=mode, chunk_frames=chunk)
        y1, e1 = cut.forward_wav(wav)
        d = (e0 - e1).abs().amax(dim=2).cpu().numpy()
        print(mode, "chunk", chunk, "max diff per (clip, frame):")
        print(np.array2string(d, precision=2))
