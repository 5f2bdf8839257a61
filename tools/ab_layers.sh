for lib in libflownet_base.so libflownet.so; do
  FLOWNET_LIB=$PWD/steady-state-flow-with-neural-nets_b200/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --layers-out gpurun_out/layers_$lib.json > /dev/null 2>&1
done
