"""Small runs of every kernel family for compute-sanitizer (memcheck / racecheck): offline whole and time-chunked (carry
arena, slot-table fallback with history rows, fused DDB block with chunk history), streaming incl. graph capture / replay,
the fused tail kernel (knob), and the int8-hybrid variant."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nunet_b200.engine import NunetEngine
from nunet_b200.synth import synth_clips
from nunet_b200.weights import load_default_weights, pack_blob, random_ddb_weights, random_lstm_weights

fuse = os.environ.get("NUNET_STREAM_FUSE", "0")
for variant, w in ((0, random_lstm_weights(1)), (1, random_ddb_weights(1))):
    B, T = 3, 40    # long enough that most tiles lie inside one clip (tensor-map boxes, CTA pairs)
    wav = torch.from_numpy(synth_clips(B, 512 + 256 * (T - 1))).cuda()
    eng = NunetEngine(pack_blob(w, variant), max_frames=B * T, max_streams=2, variant=variant)
    y, est = eng.forward_wav(wav)
    cut = NunetEngine(pack_blob(w, variant), max_frames=B * T, variant=variant, chunk_frames=13)      # chunks of 13, 13, 13, 1
    y2, est2 = cut.forward_wav(wav)
    assert torch.equal(est[:, :39], est2[:, :39])
    small = NunetEngine(pack_blob(w, variant), max_frames=25, variant=variant)                        # arena smaller than a clip
    y3, _ = small.forward_wav(wav)
    eng.stream_reset()
    for t in range(6):    # 2 eager steps, 2 graph captures, 2 replays
        eng.stream_step_wav(torch.from_numpy(synth_clips(2, 256)).cuda())
    torch.cuda.synchronize()
    print("variant", variant, "fuse", fuse, "ok", float(est.abs().max()))
    eng.close(); cut.close(); small.close()
from nunet_b200.tflite_export import hybrid_weight_set
h = NunetEngine(pack_blob(hybrid_weight_set(load_default_weights()), 2), max_streams=2, variant=2)
h.stream_reset()
for t in range(4):
    out = h.stream_step_mag(torch.rand(2, 256, device="cuda") * 5)
torch.cuda.synchronize()
print("variant 2 (int8-hybrid) ok", float(out.abs().max()))
