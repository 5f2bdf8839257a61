"""BASELINE.json configs beside the headline one: device-resident CUDA-graph throughput of flow_net.inference for
configs[0] (batch 8 synthetic), configs[3] (512x1024) and configs[4] (nr_res_blocks=2, filter_size=16, batch sweep).
Prints one JSON object per line; numbers go into BASELINE.md section 5."""
import json
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "steady-state-flow-with-neural-nets_b200"))
sys.path.insert(0, ROOT)
import flownet_native as native  # noqa: E402
import model.flow_architecture as arch  # noqa: E402
import model.flow_net as flow_net  # noqa: E402
from model.flags import FLAGS  # noqa: E402
from utils.synthetic import synth_boundary  # noqa: E402


def run(name, B, H, W, steps=30, gflop_per_img=2.841, **flags):
    old = {k: getattr(FLAGS, k) for k in flags}
    for k, v in flags.items():
        setattr(FLAGS, k, v)
    try:
        arch.reset_variables(1234)
        x = torch.from_numpy(synth_boundary(B, H, W, seed=0)).cuda()
        net = flow_net.CompiledInference(B, (H, W))
        net(x)
        for _ in range(3):
            net.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            net.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        ips = B / (ms * 1e-3)
        print(json.dumps({"config": name, "batch": B, "shape": [H, W], "flags": flags, "ms_per_step": round(ms, 4),
                          "images_per_s": round(ips, 1), "tflops": round(ips * gflop_per_img / 1e3, 1),
                          "launches": net.launches_per_replay, "tc_status": native.tc_status()}), flush=True)
        del net
    finally:
        for k, v in old.items():
            setattr(FLAGS, k, v)
        arch.reset_variables(1234)
        torch.cuda.empty_cache()


run("configs[0] batch 8 synthetic", 8, 128, 256)
run("configs[1]-like batch 64 synthetic", 64, 128, 256)
run("batch 256 synthetic", 256, 128, 256, steps=10)
run("configs[3] 512x1024 batch 4 (one GPU's share)", 4, 512, 1024, gflop_per_img=45.46)
run("configs[3] 512x1024 batch 32", 32, 512, 1024, steps=5, gflop_per_img=45.46)
for b in (1, 8, 64, 256):
    run("configs[4] nr_res_blocks=2 filter_size=16", b, 128, 256, steps=10 if b >= 64 else 30, gflop_per_img=19.49,
        nr_res_blocks=2, filter_size=16)
