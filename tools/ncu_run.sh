set -u
cd /root/repo
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/rp_traffic.csv python tools/ncu_target.py 256 > gpurun_out/rp_traffic.log 2>&1
NCU_BATCH=64 timeout 300 bash tools/ncu_capture.sh rp_spconv6 conv_tc3 12 1 source
NCU_BATCH=64 timeout 300 bash tools/ncu_capture.sh rp_deconv1 conv_tc3 110 1
NCU_BATCH=64 timeout 300 bash tools/ncu_capture.sh rp_enin conv_tc3 0 1
NCU_BATCH=64 timeout 300 bash tools/ncu_capture.sh rp_dein conv_tc3 109 1
NUNET_DEBUG_KNOBS=1 NUNET_TC3_TIMING=1 timeout 200 python tools/ncu_target.py 256 2>&1 | grep TC3TIMING > gpurun_out/rp_role_cycles.txt
rm -f gpurun_out/rp_*.ncu-rep
wc -l gpurun_out/rp_traffic.csv gpurun_out/rp_role_cycles.txt
