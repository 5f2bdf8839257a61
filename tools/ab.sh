# A/B two builds of libflownet.so on the same box: base (previous commit) vs new (working tree)
for rep in 1 2; do
for lib in libflownet_base.so libflownet.so; do
  echo -n "$lib: "
  FLOWNET_LIB=$PWD/steady-state-flow-with-neural-nets_b200/$lib timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | python -c "import json,sys; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(b['value']), b['ms_per_step'], b['clocks'])"
done
done
