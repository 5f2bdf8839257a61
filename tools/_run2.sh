MMA_BENCH3=1 timeout 120 tools/mma_bench > gpurun_out/r2_mma_bench3.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -o /tmp/r2_fwd_full -f python tools/ncu_forward.py > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/r2_fwd_full.ncu-rep --page raw --csv > gpurun_out/r2_fwd_full_raw.csv 2>/dev/null
cat gpurun_out/r2_mma_bench3.log
