timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "level4 or level5 or net_vs or golden or smoke" 2>&1 | tail -5
timeout 300 python bench.py --no-train --no-cpu-baseline --layers-out gpurun_out/r2_layers_v17.json > gpurun_out/r2_bench_v17.json 2> gpurun_out/r2_bench_v17.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_v17.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['launch'])"
timeout 300 python bench.py --no-train --no-cpu-baseline --batch 8 > gpurun_out/r2_bench_v17_b8.json 2>> gpurun_out/r2_bench_v17.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_v17_b8.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
tail -3 gpurun_out/r2_bench_v17.err
