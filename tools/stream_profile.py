"""Per-launch CUDA-event times of one streaming step (S streams), aggregated by kernel kind and by block.
usage: python tools/stream_profile.py [S]"""
import collections, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nunet_b200.engine import NunetEngine
from nunet_b200.synth import synth_clips
from nunet_b200.weights import load_default_weights, pack_blob
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
eng = NunetEngine(pack_blob(load_default_weights()), max_streams=S, ctfa_mode="frame_div32")
hops = torch.from_numpy(np.tile(synth_clips(32, 256 * 8), (S // 32 + 1, 1))[:S]).cuda()
out = torch.empty((S, 256), device="cuda")
eng.stream_reset()
for t in range(4):
    eng.stream_step_wav(hops[:, 256 * t:256 * (t + 1)].contiguous(), out)
eng.profile(True)
eng.stream_step_wav(hops[:, 256 * 4:256 * 5].contiguous(), out)
torch.cuda.synchronize()
ent = eng.profile_entries()
eng.profile(False)
tot = sum(e[1] for e in ent)
print(f"S={S}: {len(ent)} launches, {tot:.3f} ms (eager, event after every launch)")
kind = collections.defaultdict(lambda: [0.0, 0, 0.0])
blk = collections.defaultdict(lambda: [0.0, 0])
for n, ms, b in ent:
    role, k = (n.split(":") + [""])[:2]
    kind[k][0] += ms; kind[k][1] += 1; kind[k][2] += b
    blk[role.split("_conv")[0].split("_spconv")[0].split("_in")[0].split("_bb")[0]][0] += ms
for k, (ms, c, b) in sorted(kind.items(), key=lambda x: -x[1][0]):
    print(f"  {k:24s} {c:4d} launches {ms:7.3f} ms  avg {1e3 * ms / c:6.1f} us   alg {b / 1e9:6.3f} GB")
print("slowest:", [(n, round(1e3 * ms, 1)) for n, ms, b in sorted(ent, key=lambda e: -e[1])[:12]])
print("fastest conv:", [(n, round(1e3 * ms, 1)) for n, ms, b in sorted([e for e in ent if 'conv_tc3' in e[0]], key=lambda e: e[1])[:8]])
