"""BASELINE.json configs beside the headline one: device-resident CUDA-graph throughput of flow_net.inference for
configs[0] (batch 8 synthetic), configs[3] (512x1024, batch 32) and configs[4] (nr_res_blocks=2, filter_size=16, batch
sweep 1-1024).  One process per GPU (run it under torch.distributed.run for N > 1): the GLOBAL batch is sharded over the
ranks (images are independent: no collective), every rank times its shard with CUDA events behind a barrier, the slowest
rank counts.  Prints one JSON object per line with the per-GPU fraction of the measured sustained bf16 peak
(MEASURED_PEAKS.json); the committed outputs are profiles/r2_config_sweep_{1,8}gpu.jsonl."""
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


RANK, WORLD, LOCAL = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(LOCAL)
import torch.distributed as dist  # noqa: E402
if WORLD > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", LOCAL))
try:
    TC_SUSTAINED = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
except Exception:
    TC_SUSTAINED = None


def run(name, GB, H, W, steps=30, gflop_per_img=2.841, **flags):
    """GB: global batch; this rank takes GB / WORLD images (ranks beyond a small batch idle)."""
    B = GB // WORLD if GB >= WORLD else (1 if RANK < GB else 0)
    if B == 0:          # nothing for this rank: take part in the barriers only
        if WORLD > 1:
            dist.barrier(); dist.barrier()
            t = torch.zeros(1, dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return
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
        if WORLD > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            net.replay()
        e1.record()
        torch.cuda.synchronize()
        if WORLD > 1:
            dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device="cuda")
        if WORLD > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        ips = min(GB, B * WORLD) / (ms * 1e-3)
        if RANK == 0:
            tf = ips * gflop_per_img / 1e3
            print(json.dumps({"config": name, "n_gpus": WORLD, "global_batch": GB, "batch_per_gpu": B, "shape": [H, W], "flags": flags,
                              "ms_per_step": round(ms, 4), "images_per_s": round(ips, 1), "tflops": round(tf, 1),
                              "frac_of_sustained_bf16_per_gpu": round(tf / min(WORLD, GB) / TC_SUSTAINED, 3) if TC_SUSTAINED else None,
                              "launches": net.launches_per_replay, "tc_status": native.tc_status()}), flush=True)
        del net
    finally:
        for k, v in old.items():
            setattr(FLAGS, k, v)
        arch.reset_variables(1234)
        torch.cuda.empty_cache()


run("configs[0] batch 8 synthetic", 8, 128, 256)
run("configs[1]-like batch 64 per GPU synthetic", 64 * WORLD, 128, 256)
run("batch 256 per GPU synthetic", 256 * WORLD, 128, 256, steps=10)
run("configs[3] 512x1024 batch 32", 32, 512, 1024, steps=10, gflop_per_img=45.46)
if WORLD > 1:
    run("configs[3] 512x1024 batch 32 per GPU", 32 * WORLD, 512, 1024, steps=5, gflop_per_img=45.46)
for gb in (1, 8, 64, 256, 1024):
    run("configs[4] nr_res_blocks=2 filter_size=16", gb, 128, 256, steps=10 if gb // WORLD >= 64 else 30, gflop_per_img=19.49,
        nr_res_blocks=2, filter_size=16)
if WORLD > 1:
    dist.barrier()
    dist.destroy_process_group()
