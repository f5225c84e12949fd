#!/usr/bin/env bash
# Runs ON THE GPU BOX (gpurun -- 'bash tools/refresh_profiles.sh'): regenerates everything profiles/ is built from into
# gpurun_out/.  Post-process here with tools/profile_classes.py, tools/ncu_traffic.py and tools/ncu_full_table.py.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/rp_tests.log 2>&1; tail -n 1 gpurun_out/rp_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/rp_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/rp_bench.json 2> gpurun_out/rp_bench.err
cp gpurun_out/bench_kernel_profile_lstm.json gpurun_out/rp_kp_lstm.json; cp gpurun_out/bench_kernel_profile_ddb.json gpurun_out/rp_kp_ddb.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/rp_ref_bench.json 2>> gpurun_out/rp_bench.err
# ncu launch list of the same command (per-launch times are cold-cache and serialised: shares must agree, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/rp_launches.csv \
    python bench.py --config offline --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/rp_ncu_bench.log 2>&1
# DRAM bytes per launch of one full-size forward (256 clips x 249 frames)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/rp_traffic.csv python tools/ncu_target.py 256 > gpurun_out/rp_traffic.log 2>&1
# ncu --set full of the four heaviest conv units (batch 64) and of the LSTM kernel
NCU_BATCH=64 timeout 300 bash tools/ncu_capture.sh rp_spconv6 conv_tc3 12 1 source
NCU_BATCH=64 timeout 300 bash tools/ncu_capture.sh rp_deconv1 conv_tc3 110 1
NCU_BATCH=64 timeout 300 bash tools/ncu_capture.sh rp_enin conv_tc3 0 1
NCU_BATCH=64 timeout 300 bash tools/ncu_capture.sh rp_dein conv_tc3 109 1
NCU_BATCH=64 timeout 300 bash tools/ncu_capture.sh rp_lstm lstm_block 0 1
NUNET_DEBUG_KNOBS=1 NUNET_TC3_TIMING=1 timeout 200 python tools/ncu_target.py 256 2>&1 | grep TC3TIMING > gpurun_out/rp_role_cycles.txt
timeout 200 python tools/stream_profile.py 1024 > gpurun_out/rp_stream_profile.txt 2>&1
rm -f gpurun_out/rp_*.ncu-rep
cut -c1-300 gpurun_out/rp_bench.json; cut -c1-200 gpurun_out/rp_ref_bench.json
