"""Data-parallel training check on real GPUs: N ranks x (B/N) images with the two-bucket overlapped all-reduce
(flow_net.CompiledTraining) must follow the single-GPU full-batch run.
  python tools/check_ddp.py --save /tmp/ref.pt                      (one GPU, global batch)
  torchrun --nproc-per-node 2 tools/check_ddp.py --compare /tmp/ref.pt
Prints the relative L2 distance of the parameters after --steps steps and whether all ranks hold identical parameters."""
import argparse, os, sys
import numpy as np
import torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "steady-state-flow-with-neural-nets_b200"))
ap = argparse.ArgumentParser()
ap.add_argument("--save")
ap.add_argument("--compare")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--batch", type=int, default=8)
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
import torch.distributed as dist
if world > 1:
    dist.init_process_group("nccl")
import model.flow_architecture as arch
import model.flow_net as flow_net
from oracle import flow_oracle as fo
B, H, W = args.batch, 128, 256
x = fo.synth_boundary(B, H, W, seed=3).astype(np.uint8)
rng = np.random.default_rng(4)
tgt = (0.2 * np.tanh(rng.standard_normal((B, H, W, 2))) * (1 - x)).astype(np.float32)
per = B // world
sl = slice(rank * per, (rank + 1) * per)
arch.reset_variables(1234)
flow_net.reset_training_state()
ct = flow_net.CompiledTraining(per, (H, W))
xb, tb = torch.from_numpy(x[sl]).cuda(), torch.from_numpy(tgt[sl]).cuda()
losses = []
for _ in range(args.steps):
    losses.append(ct.step(xb, tb, 1e-3))
torch.cuda.synchronize()
param = flow_net.training_state().param.clone()
loss = torch.stack([l.reshape(()) for l in losses])
if world > 1:
    dist.all_reduce(loss, op=dist.ReduceOp.SUM)
    ps = [torch.empty_like(param) for _ in range(world)]
    dist.all_gather(ps, param)
    same = all(torch.equal(ps[0], q) for q in ps[1:])
else:
    same = True
if rank == 0:
    print("graphs:", "A+B, cut at flat offset %d of %d" % (ct.bucket_offset, param.numel()) if ct.graph_b is not None else "one")
    print("losses (summed over ranks):", [round(v, 3) for v in loss.tolist()])
    print("all ranks identical:", same)
    if args.save:
        torch.save({"param": param.cpu(), "loss": loss.cpu()}, args.save)
    if args.compare:
        ref = torch.load(args.compare)
        d = (param.cpu().double() - ref["param"].double()).norm() / ref["param"].double().norm()
        init = None
        print("parameters vs the single-GPU full-batch run: rel L2 %.3e; losses rel %.3e" % (
            d.item(), ((loss.cpu() - ref["loss"]).abs() / ref["loss"].abs()).max().item()))
        assert same and d.item() < 1e-4, "data-parallel run diverges from the full-batch run"
        print("OK")
if world > 1:
    dist.destroy_process_group()
