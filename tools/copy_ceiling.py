"""Copy-only ceiling of the serving loop: what images/s the host <-> device copies alone allow at N ranks, without any
compute.  Every rank moves what one bench step moves -- a pinned uint8 batch up (2 MB at batch 64) and
a float32 field down (16.8 MB) -- on two copy streams, concurrently, for --steps steps; rank 0 prints per-rank and
aggregate GB/s and the images/s they would sustain.  Run it the way bench.py is run:

  python tools/copy_ceiling.py                                   (1 GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/copy_ceiling.py

--no-numa leaves the process unpinned (the round-1 configuration)."""
import argparse, json, os, sys
import torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "steady-state-flow-with-neural-nets_b200"))
from utils.numa import bind_to_gpu_numa_node

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--no-numa", action="store_true")
args = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
import torch.distributed as dist
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
numa = None if args.no_numa else bind_to_gpu_numa_node(local)
B = args.batch
x_host = torch.zeros((B, 128, 256, 1), dtype=torch.uint8).pin_memory()
y_host = [torch.empty((B, 128, 256, 2), dtype=torch.float32).pin_memory() for _ in range(2)]
x_dev = [torch.empty_like(x_host, device="cuda") for _ in range(2)]
y_dev = [torch.zeros((B, 128, 256, 2), dtype=torch.float32, device="cuda") for _ in range(2)]
up, down = torch.cuda.Stream(), torch.cuda.Stream()


def loop(n):
    for i in range(n):
        with torch.cuda.stream(up):
            x_dev[i & 1].copy_(x_host, non_blocking=True)
        with torch.cuda.stream(down):
            y_host[i & 1].copy_(y_dev[i & 1], non_blocking=True)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


loop(10)
barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
loop(args.steps)
up.synchronize(); down.synchronize()
e1.record()
barrier()
ms = e0.elapsed_time(e1)
t = torch.tensor([ms], dtype=torch.float64, device="cuda")
allms = [torch.zeros_like(t) for _ in range(world)]
if world > 1:
    dist.all_gather(allms, t)
else:
    allms = [t]
if rank == 0:
    per_rank = [float(a.item()) for a in allms]
    worst = max(per_rank)
    d2h = B * 128 * 256 * 2 * 4
    h2d = B * 128 * 256
    print(json.dumps({
        "n_gpus": world, "batch_per_gpu": B, "steps": args.steps, "numa": numa, "ms_per_step_per_rank": [m / args.steps for m in per_rank],
        "d2h_GBps_per_rank": [d2h * args.steps / (m * 1e-3) / 1e9 for m in per_rank],
        "aggregate_GBps": (d2h + h2d) * args.steps * world / (worst * 1e-3) / 1e9,
        "copy_only_images_per_s": B * args.steps * world / (worst * 1e-3)}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
