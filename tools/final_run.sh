set -u
cd /root/repo
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/rp_tests.log 2>&1; tail -n 1 gpurun_out/rp_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py > gpurun_out/rp_bench.json 2> gpurun_out/rp_bench.err; cp gpurun_out/bench_kernel_profile_lstm.json gpurun_out/rp_kp_lstm.json
timeout 300 python bench.py --config streaming --no-cpu-baseline > gpurun_out/rp_stream_bench.json 2>> gpurun_out/rp_bench.err
timeout 300 python bench.py --variant ddb --no-cpu-baseline > gpurun_out/rp_ddb_bench.json 2>> gpurun_out/rp_bench.err; cp gpurun_out/bench_kernel_profile_lstm.json gpurun_out/rp_kp_ddb.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/rp_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/rp_ncu_bench.log 2>&1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/rp_ref_bench.json 2>> gpurun_out/rp_bench.err
cut -c1-200 gpurun_out/rp_bench.json; cut -c1-160 gpurun_out/rp_stream_bench.json; cut -c1-160 gpurun_out/rp_ddb_bench.json; cut -c1-200 gpurun_out/rp_ref_bench.json
