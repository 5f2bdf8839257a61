timeout 300 python bench.py --no-train --no-cpu-baseline --layers-out gpurun_out/r2_layers_v21.json > gpurun_out/r2_bench_v21.json 2> gpurun_out/r2_bench_v21.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_v21.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['launch'])"
timeout 300 python bench.py --no-train --no-cpu-baseline --batch 8 > gpurun_out/r2_bench_v21_b8.json 2>> gpurun_out/r2_bench_v21.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_v21_b8.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['launch'])"
tail -3 gpurun_out/r2_bench_v21.err
