#!/usr/bin/env bash
# One GPU visit: parity tests, the default bench line (all three configs), per-class kernel times.  Outputs under gpurun_out/<tag>_*.
tag=${1:-r2}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${tag}_tests.log
tail -25 gpurun_out/${tag}_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/${tag}_bench.json
python tools/profile_classes.py gpurun_out/bench_kernel_profile_lstm.json 12 > gpurun_out/${tag}_classes_lstm.txt 2>&1
python tools/profile_classes.py gpurun_out/bench_kernel_profile_ddb.json 8 > gpurun_out/${tag}_classes_ddb.txt 2>&1
cat gpurun_out/${tag}_classes_lstm.txt
