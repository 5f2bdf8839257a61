timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "wgrad or train or bias or backward or grad or adam or dropout" 2>&1 | tail -3
for v in 0 1 1; do FLOWNET_TRAIN_STREAMS=$v timeout 300 python bench.py --mode train --no-cpu-baseline --steps 30 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('train streams=$v', d['value'], d['ms_per_step'], d['gpu_launches'], d.get('loss'))"; done
for v in 0 1; do FLOWNET_TRAIN_STREAMS=$v timeout 300 python bench.py --mode train --no-cpu-baseline --train-batch 32 --steps 30 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('train b32 streams=$v', d['value'], d['ms_per_step'], d['gpu_launches'], d.get('loss'))"; done
