timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tail or golden or net_vs or smoke" 2>&1 | tail -3
for i in 1 2; do timeout 300 python bench.py --no-train --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['launch'])"; done
