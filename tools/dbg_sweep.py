"""Experiment driver: per-class conv kernel times with parts of conv_tc3 disabled (NUNET_TC3_DBG bitmask)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
names = ["msfe6_de_conv1", "msfe6_en_conv1", "msfe6_en_conv2", "msfe6_de_in", "msfe6_en_spconv6", "msfe6_en_spconv5", "msfe6_en_in",
         "msfe6_down_sampling", "msfe5_de_conv1", "msfe4_en_spconv4"]
rows = {}
for dbg in sys.argv[1:]:
    env = dict(os.environ, NUNET_DEBUG_KNOBS="1", NUNET_TC3_DBG=dbg)
    subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu-baseline", "--steps", "2"], env=env,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=600)
    d = json.load(open(os.path.join(ROOT, "gpurun_out", "bench_kernel_profile_lstm.json")))
    t = {n.split(":")[0]: ms for n, ms, b in d["entries"]}
    rows[dbg] = (d["total_ms"], [t.get(n, 0.0) for n in names])
print("dbg    total  " + " ".join(f"{n[-12:]:>12s}" for n in names))
for dbg, (tot, v) in rows.items():
    print(f"{dbg:>3s} {tot:8.2f}  " + " ".join(f"{x:12.3f}" for x in v))
