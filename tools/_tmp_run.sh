timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "level or net_ or golden" 2>&1 | tail -4
for i in 1 2; do timeout 300 python bench.py --no-train --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['launch'], d['roofline']['kernel'][:30], d['roofline']['kernel_ms'], d['roofline']['frac'])"; done
