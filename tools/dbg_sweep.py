"""Experiment driver: per-class conv kernel times with parts of conv_tc3 disabled (NUNET_TC3_DBG bitmask)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for dbg in sys.argv[1:]:
    env = dict(os.environ, NUNET_TC3_DBG=dbg)
    subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu-baseline", "--steps", "2"], env=env,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=600)
    print("== dbg", dbg, flush=True)
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "profile_classes.py"),
                    os.path.join(ROOT, "gpurun_out", "bench_kernel_profile.json")])
