timeout 300 python tools/phase_timing_innet.py 2>&1 | tail -4
bash tools/run_train_smoke.sh 2>&1 | tail -24
