timeout 300 python tools/phase_timing_innet.py 2>&1 | tail -5
