"""world_size-2 gloo tests (CPU) of the multi-process logic: the data-parallel gradient exchange of the
training step (model/flow_backward.allreduce_gradients: SUM, no 1/R -- SURVEY 8e) checked against the
full-batch gradient of the CPU oracle, and bench.py's weak-scaling aggregation (sum of images, max of
times).  The CUDA kernels themselves cannot run here; their parity tests are the -m gpu suite."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "steady-state-flow-with-neural-nets_b200")


def _flat_oracle_grad(x, tgt):
    from oracle import flow_oracle as fo
    vt = fo.VariableStore(1234, torch.float64)
    fo.inference(vt, x, 1.0)
    vt.vars = {k: v.clone().requires_grad_(True) for k, v in vt.vars.items()}
    fo.loss_image(fo.inference(vt, x, 1.0), tgt).backward()
    return torch.cat([v.grad.reshape(-1) for v in vt.vars.values()])


def _worker(rank, world, port, out_dir):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import flow_oracle as fo
        import model.flow_backward as fb           # imports the ctypes binding; no CUDA call is made
        B = 4
        x = torch.from_numpy(fo.synth_boundary(B, 16, 32, seed=7)).double()
        rng = np.random.default_rng(5)
        tgt = torch.from_numpy(0.1 * np.tanh(rng.standard_normal((B, 16, 32, 2)))) * (1 - x)
        per = B // world
        shard = slice(rank * per, (rank + 1) * per)
        g = _flat_oracle_grad(x[shard], tgt[shard])
        # the two-bucket exchange of flow_net.CompiledTraining (decoder + deepest level first, encoder second) must equal
        # the single all-reduce: cut from flow_backward.bucket_split on the oracle's variable layout
        vs = fo.VariableStore(1234, torch.float64)
        fo.inference(vs, x[:1], 1.0)

        class _Flat:
            offsets = {}
        off = 0
        for name, v in vs.vars.items():
            _Flat.offsets[name] = (off, v.numel(), tuple(v.shape))
            off += v.numel()
        blocks = sorted({n.split("_conv_1_conv/")[0] for n in vs.vars if "_conv_1_conv/weights" in n},
                        key=lambda b: _Flat.offsets[b + "_conv_1_conv/weights"][0])
        tape = [dict(kind="block", name=b) for b in blocks] + [dict(kind="last")]
        k, cut = fb.bucket_split(tape, _Flat)
        assert tape[k]["name"] == "resnet_5_downsample" and 0 < cut < g.numel()
        assert cut == sum(v.numel() for n, v in vs.vars.items() if _Flat.offsets[n][0] < _Flat.offsets["resnet_5_downsample_conv_1_conv/weights"][0])
        g2 = g.clone()
        dist.all_reduce(g2[cut:], op=dist.ReduceOp.SUM)
        dist.all_reduce(g2[:cut], op=dist.ReduceOp.SUM)
        g = fb.allreduce_gradients(g)
        assert torch.equal(g, g2)
        full = _flat_oracle_grad(x, tgt)
        err = ((g - full).norm() / full.norm()).item()
        # bench.py aggregation: every rank times its own steps, value = all images / max time
        t = torch.tensor([1.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        np.save(os.path.join(out_dir, f"r{rank}.npy"), np.array([err, t.item(), g.numel()]))
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_full_batch(tmp_path, built_lib):
    port = 29650 + (os.getpid() % 200)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        err, tmax, n = np.load(tmp_path / f"r{r}.npy")
        assert err < 1e-12            # fp64: shard gradients of a SUM loss add exactly (up to summation order)
        assert tmax == 2.0            # max over ranks
        assert int(n) == 2_561_386
