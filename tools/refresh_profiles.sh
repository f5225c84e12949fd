#!/usr/bin/env bash
# Runs ON THE GPU BOX (gpurun -- 'bash tools/refresh_profiles.sh'): regenerates everything profiles/ is built from into
# gpurun_out/.  Post-process here with tools/profile_classes.py, tools/ncu_traffic.py and tools/ncu_summary.py.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/rp_tests.log 2>&1; tail -n 1 gpurun_out/rp_tests.log
timeout 300 python bench.py > gpurun_out/rp_bench.json 2> gpurun_out/rp_bench.err; cp gpurun_out/bench_kernel_profile_lstm.json gpurun_out/rp_kp_lstm.json
timeout 300 python bench.py --config streaming --no-cpu-baseline > gpurun_out/rp_stream_bench.json 2>> gpurun_out/rp_bench.err
timeout 300 python bench.py --variant ddb --no-cpu-baseline > gpurun_out/rp_ddb_bench.json 2>> gpurun_out/rp_bench.err; cp gpurun_out/bench_kernel_profile_lstm.json gpurun_out/rp_kp_ddb.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/rp_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/rp_ncu_bench.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/rp_traffic.csv python tools/ncu_target.py 256 > gpurun_out/rp_traffic.log 2>&1
NCU_BATCH=64 timeout 300 bash tools/ncu_capture.sh rp_spconv6 conv_tc3 12 1 source
NCU_BATCH=64 timeout 300 bash tools/ncu_capture.sh rp_deconv1 conv_tc3 110 1
NCU_BATCH=64 timeout 300 bash tools/ncu_capture.sh rp_enin conv_tc3 0 1
NCU_BATCH=64 timeout 300 bash tools/ncu_capture.sh rp_dein conv_tc3 109 1
NUNET_DEBUG_KNOBS=1 NUNET_TC3_TIMING=1 timeout 200 python tools/ncu_target.py 256 2>&1 | grep TC3TIMING > gpurun_out/rp_role_cycles.txt
rm -f gpurun_out/rp_*.ncu-rep
cut -c1-200 gpurun_out/rp_bench.json; cut -c1-160 gpurun_out/rp_stream_bench.json; cut -c1-160 gpurun_out/rp_ddb_bench.json
