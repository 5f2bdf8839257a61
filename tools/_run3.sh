timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "level5 or level4" 2>&1 | tail -3
timeout 300 python tools/phase_timing_l5.py > gpurun_out/l5_phase.log 2>&1; head -4 gpurun_out/l5_phase.log
AB_BATCHES=64 timeout 200 python tools/ab_level5.py 2>&1 | tail -2
