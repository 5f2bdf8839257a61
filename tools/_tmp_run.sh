timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "train or grad or backward or boundary" 2>&1 | tail -3
for i in 1 2; do timeout 300 python bench.py --mode train --no-cpu-baseline --steps 30 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('train', d['value'], d['ms_per_step'], d['gpu_launches'])"; done
