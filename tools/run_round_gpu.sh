#!/bin/bash
# One GPU-box pass that produces the evidence committed under profiles/: bench line (+ per-layer table, train sub-record),
# `ncu --set full` of one warm forward (raw page as CSV; the .ncu-rep itself stays on the box), ncu launch list of the bench.
set -x
mkdir -p gpurun_out
if [ "$1" = "tests" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
fi
timeout 600 python bench.py --layers-out gpurun_out/r2_layers_final.json > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -o /tmp/r2_fwd_full -f python tools/ncu_forward.py > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/r2_fwd_full.ncu-rep --page raw --csv > gpurun_out/r2_fwd_full_raw.csv 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_train.csv python bench.py --mode train --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_train.log 2>&1
timeout 300 python bench.py --mode train --no-cpu-baseline --train-batch 32 --steps 30 > gpurun_out/r2_bench_train_b32.json 2>/dev/null
tail -3 gpurun_out/r2_pytest.log
head -c 600 gpurun_out/r2_bench_final.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
