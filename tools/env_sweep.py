"""Experiment driver: bench totals and selected layer times under different environment settings.
usage: python tools/env_sweep.py "A=1 B=2" "A=3" ..."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
names = ["msfe6_de_conv1", "msfe6_en_conv1", "msfe6_en_conv2", "msfe6_de_in", "msfe6_en_spconv6", "msfe6_en_spconv5", "msfe6_en_in",
         "msfe6_down_sampling", "msfe5_de_conv1", "msfe4_en_spconv4"]
print("                  total     " + " ".join(f"{n[-12:]:>12s}" for n in names))
for spec in sys.argv[1:]:
    env = dict(os.environ)
    for kv in spec.split():
        k, v = kv.split("=")
        env[k] = v
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu-baseline", "--steps", "4"], env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=600, text=True)
    try:
        step_ms = json.loads(r.stdout.strip().splitlines()[-1])["ms_per_step"]
    except Exception:
        step_ms = float("nan")
    d = json.load(open(os.path.join(ROOT, "gpurun_out", "bench_kernel_profile_lstm.json")))
    t = {n.split(":")[0]: ms for n, ms, b in d["entries"]}
    print(f"step {step_ms:7.2f} ms | {d['total_ms']:8.2f}  " + " ".join(f"{t.get(n, 0.0):12.3f}" for n in names) + "   " + spec, flush=True)
