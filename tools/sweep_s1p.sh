for kb in 112 70 50 30; do
  echo "== smem cap $kb KB"
  FLOWNET_S1P_SMEM_KB=$kb timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --layers-out gpurun_out/layers_cap$kb.json 2>&1 | python -c "import json,sys; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(b['value'], b['ms_per_step'])"
done
