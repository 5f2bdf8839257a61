"""Per-layer times of the wider/deeper variant (nr_res_blocks=2, filter_size=16) -- which layers still miss a
specialised kernel."""
import os, sys, json
import torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "steady-state-flow-with-neural-nets_b200")); sys.path.insert(0, ROOT)
import bench
import model.flow_architecture as arch
from model.flags import FLAGS
from utils.synthetic import synth_boundary
FLAGS.nr_res_blocks = int(os.environ.get("NR", "2")); FLAGS.filter_size = int(os.environ.get("FS", "16"))
B = int(os.environ.get("B", "64"))
arch.reset_variables(1234)
x = torch.from_numpy(synth_boundary(B, 128, 256, seed=0)).cuda()
rows = bench.layer_profile(x, reps=3)
tot = sum(r["ms"] for r in rows)
for r in rows:
    print("%-36s %-5s %-10s K %4d N %4d s%d tc=%d %8.1fus %6.1fTF" % (r["name"], r["kind"], r["out_hw"], r["K_tap"], r["N"], r["stride"],
          int(r.get("tcgen05", 0)), r["ms"] * 1e3, r["flops"] / r["ms"] / 1e9))
print("sum ms", tot)
