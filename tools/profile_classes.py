"""Aggregate bench_kernel_profile_lstm.json (per-launch CUDA-event times from bench.py) by unit class."""
import collections
import json
import re
import sys

d = json.load(open(sys.argv[1]))
agg = collections.defaultdict(lambda: [0.0, 0.0, 0])
for n, ms, b in d["entries"]:
    role, k = n.split(":")
    if not k.startswith("conv"):
        cls = k
    elif "_spconv" in role:
        last = re.search(r"msfe6_\w+_spconv6|msfe5_\w+_spconv5|msfe4_\w+_spconv4|msfe3_\w+_spconv3", role)
        cls = "spconv_last(N128)" if last else "spconv(N64)"
    elif "_conv" in role:
        cls = "conv_s2(N32)"
    elif "down" in role:
        cls = "down(N64)"
    elif role.endswith("_in"):
        cls = "in_dec(up+in,N128)" if "_de" in role else "in_enc(1x1,N64)"
    else:
        cls = role
    agg[cls][0] += ms
    agg[cls][1] += b
    agg[cls][2] += 1
print(f"total {d['total_ms']:.2f} ms")
for k, (ms, b, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
    print(f"{k:22s} {c:4d} launches {ms:8.2f} ms {100 * ms / d['total_ms']:5.1f}%  alg {b / 1e9:7.2f} GB -> {b / ms / 1e6:7.0f} GB/s"
          f"  hbm-floor {b / 6.5434e9:6.2f} ms")
if len(sys.argv) > 2:
    for n, ms, b in sorted(d["entries"], key=lambda e: -e[1])[: int(sys.argv[2])]:
        print(f"{n:40s} {ms:7.3f} ms  {b / ms / 1e6:8.1f} GB/s")
