#!/usr/bin/env python
"""bench.py -- flow_net images/sec at 128x256 on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one forward pass of flow_net (default flags, seed-1234 Xavier weights) over one batch
of 64 boundary images per GPU: BASELINE configs[1], the 28 bundled car masks tiled to 64
(fixture tests/golden/cars_128x256.npz).  Multi-GPU shards by batch (weak scaling, no collective).

  value     device-resident throughput: inputs already in HBM, one CUDA-graph replay per step
  e2e       same metric through the public API with HOST buffers: pinned uint8 boundaries H2D,
            forward, fp32 velocity field D2H, every step, inside the timed region
  roofline  the dominant kernel of the step, timed alone with CUDA events (per-launch)
  cpu_baseline  the CPU oracle (torch fp32, all host threads) on a bounded sample (rank 0, N=1)

--impl reference times the CPU oracle port instead (the TF reference cannot run here: SURVEY F11).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "steady-state-flow-with-neural-nets_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "flow_net images/sec at 128x256"
UNIT = "images/s"
BATCH = 64
SHAPE = (128, 256)
FLOP_PER_IMG = 2.841e9        # 1420.56 MMAC forward
BYTES_PER_IMG = 15.76e6       # bf16 activations, layer-fused, weights excluded
PUBLISHED_IMG_S = 1.0 / 0.00287   # reference README.md:42 (GTX 1080, batch 8, incl. feed/fetch)
WORKLOAD = ("configs[1]: flow_net inference, default flags, 28 bundled car boundaries tiled to batch 64 per GPU, "
            "128x256, seed-1234 Xavier weights")


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tc_burst=p["bf16_tflops"], tc_sustained=p["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, source="fallback")


def car_batch(n):
    from utils.synthetic import load_car_boundaries
    cars = load_car_boundaries(os.path.join(ROOT, "tests", "golden", "cars_128x256.npz"))
    return cars[[i % 28 for i in range(n)]]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_oracle_rate(seconds_budget=12.0, batch=8, max_iters=40):
    """images/s of the CPU oracle (torch fp32, every host thread) on the first `batch` images of the workload."""
    from oracle import flow_oracle as fo
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    x = torch.from_numpy(car_batch(batch).astype(np.float32))
    vs = fo.VariableStore(1234, torch.float32)
    with torch.no_grad():
        fo.inference(vs, x, 1.0)          # warm-up (creates variables)
        fo.inference(vs, x, 1.0)
        t0 = time.perf_counter()
        it = 0
        while it < max_iters and (time.perf_counter() - t0 < seconds_budget or it < 3):
            fo.inference(vs, x, 1.0)
            it += 1
        dt = time.perf_counter() - t0
    return dict(value=batch * it / dt, unit=UNIT, cores=cores, kind="port",
                sample=f"{it} forwards of the first {batch} images of the workload, oracle/flow_oracle.py torch-CPU "
                       f"fp32 (oneDNN), {cores} threads; the TF reference is not installable (SURVEY F11)")


def run_reference(args, rank, world):
    """--impl reference: the CPU oracle port on the host cores; rank 0 only."""
    if rank != 0:
        return
    from oracle import flow_oracle as fo
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    b = args.batch                       # the same configuration as the repo arm: the whole 64-image batch per step
    x = torch.from_numpy(car_batch(b).astype(np.float32))
    vs = fo.VariableStore(1234, torch.float32)
    # the batch is walked in 8-image slices: the oneDNN path runs them about twice as fast as one 64-image call (measured on the
    # GPU box's host: 75 img/s whole, ~140 img/s sliced) -- the baseline gets its best configuration
    sl = 8 if b % 8 == 0 else b

    def forward():
        for i in range(0, b, sl):
            fo.inference(vs, x[i:i + sl], 1.0)

    with torch.no_grad():
        for _ in range(max(1, args.warmup)):
            forward()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            forward()
        dt = time.perf_counter() - t0
    v = b * args.steps / dt
    sample = (f"each step = one forward of the {b}-image workload batch in {sl}-image slices; oracle/flow_oracle.py "
              f"torch-CPU fp32, {cores} threads (TF-1.x reference not installable: SURVEY F11)")
    emit(({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": v / PUBLISHED_IMG_S, "dtype": "f32", "data": "bundled car masks (fixture), random-init weights",
        "config": {"workload": WORKLOAD, "batch_per_gpu": b, "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def ncu_evidence():
    """profiles/r2_ncu_summary.json: per conv launch of the default forward at batch 64, from the committed
    `ncu --set full` capture of this command (tools/ncu_summary.py writes it): DRAM bytes read + written, tensor-pipe
    and tensor-memory-pipe activity.  Read at run time, keyed by (layer, batch); absent file or key -> None."""
    path = os.path.join(ROOT, "profiles", "r2_ncu_summary.json")
    try:
        with open(path) as f:
            return json.load(f)
    except (OSError, ValueError):
        return None


def layer_profile(x_dev, reps=5):
    """Per-launch CUDA-event times of every kernel of one eager forward (median over reps)."""
    import model.flow_architecture as arch
    import model.flow_net as flow_net
    runs = []
    for _ in range(reps + 1):
        arch._PROFILE = []
        # hold the GPU for ~2 ms so that the host enqueues the whole forward (launches + events) ahead of it: each event
        # pair then brackets its kernel alone, without the host's per-call latency in between
        torch.cuda._sleep(4_000_000)
        flow_net.inference(x_dev, 1.0)
        torch.cuda.synchronize()
        runs.append(arch._PROFILE)
        arch._PROFILE = None
    runs = runs[1:]
    rows = []
    for i, rec in enumerate(runs[0]):
        ts = sorted(r[i]["events"][0].elapsed_time(r[i]["events"][1]) for r in runs)
        row = {k: v for k, v in rec.items() if k != "events"}
        row["ms"] = ts[len(ts) // 2]
        rows.append(row)
    return rows


def run_train(args, rank, world, dist, native, arch, flow_net, steps=None):
    """BASELINE configs[2]: training step (L2 loss, fwd+bwd+Adam), global batch 256 synthetic 128x256,
    batch-sharded with one NCCL all-reduce of the flat fp32 gradient.  keep_prob = 1.0 (the parity setting) unless
    --keep-prob.  Returns the record (rank 0) or None."""
    global_batch = args.train_batch
    steps = steps or args.steps
    B = global_batch // world
    arch.reset_variables(1234)
    flow_net.reset_training_state()
    boundary, sflow = flow_net.inputs(B, seed=rank * B)
    flow_net.begin_training()

    compiled = None
    if not args.eager_train:
        # forward + backward replayed from one CUDA graph; all-reduce and Adam stay eager (flow_net.CompiledTraining)
        compiled = flow_net.CompiledTraining(B, SHAPE, keep_prob=args.keep_prob, dropout_seed=1 + rank)

    def step():
        step.k += 1
        if compiled is not None:
            return compiled.step(boundary, sflow, 1e-4)
        p = flow_net.inference(boundary, args.keep_prob, dropout_seed=step.k)
        err = flow_net.loss_image_train(p, sflow)
        return flow_net.train(err, 1e-4)

    step.k = 0
    warm = max(args.warmup, 3)
    for _ in range(warm):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    n0 = native.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    flow_net.end_training()
    flow_net.reset_training_state()
    arch.reset_variables(1234)
    if rank != 0:
        return None
    v = global_batch * steps / (ms * 1e-3)
    peaks = load_peaks()
    per_gpu_tflops = v / world * 8.519e9 / 1e12         # 3x the forward's 2.841 GFLOP per image, on ONE GPU's share
    return {
        "metric": "flow_net training images/sec at 128x256 (fwd+bwd+Adam)", "value": v, "unit": UNIT, "n_gpus": world,
        "steps": steps, "warmup": warm, "ms_per_step": ms / steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic boundaries and targets",
        "config": {"workload": "configs[2]: training step, global batch 256, 128x256, L2 loss + TF1 Adam, "
                               "keep_prob %.2f; tcgen05 fwd + dgrad + wgrad; %s" % (
                                   args.keep_prob, "fwd+bwd replayed from a CUDA graph" if compiled is not None else "eager launches"),
                   "batch_per_gpu": B,
                   "parallelism": f"dp{world}, flat fp32 gradient all-reduce in two buckets, the first under the encoder's backward (%s)" % ("NCCL" if world > 1 else "single GPU: none")},
        # kernels of this library per step: those replayed from the graph + the eager ones (Adam) the counter saw
        "gpu_launches": native.launch_count() - n0 + (compiled.launches_per_replay * steps if compiled is not None else 0),
        "loss": float(loss.item()),
        "roofline": {"bound": "tensor", "achieved": per_gpu_tflops, "peak": peaks["tc_sustained"],
                     "unit": "TFLOP/s per GPU", "frac": per_gpu_tflops / peaks["tc_sustained"], "traffic": None,
                     "peak_source": peaks["source"] + " (sustained; whole step)"},
    }


_OUT = None


def emit(record):
    """the one JSON line, on the process's original stdout"""
    out = _OUT if _OUT is not None else sys.stdout
    out.write(json.dumps(record) + "\n")
    out.flush()


def main():
    # stdout carries exactly one JSON line: whatever the libraries print there (NCCL's version banner under torchrun) is
    # sent to stderr by pointing descriptor 1 at it and keeping a private copy of the original for emit()
    global _OUT
    try:
        keep = os.dup(1)
        os.dup2(2, 1)
        _OUT = os.fdopen(keep, "w")
    except OSError:              # no usable stderr: leave stdout as it is
        _OUT = None
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="images per GPU per step (default: the configs[1] batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="inference", choices=["inference", "train"],
                    help="train: only BASELINE configs[2] (fwd+bwd+Adam at global batch 256); the default run carries it as the `train` sub-record")
    ap.add_argument("--no-train", action="store_true", help="skip the `train` sub-record of the default run")
    ap.add_argument("--train-batch", type=int, default=256,
                    help="global batch of the training step (configs[2]: 256; 32 on one GPU = the per-GPU load of an 8-GPU run)")
    ap.add_argument("--train-steps", type=int, default=10, help="timed steps of the `train` sub-record")
    ap.add_argument("--keep-prob", type=float, default=1.0,
                    help="train mode: dropout keep probability (1.0 = the parity setting; the reference trains with 0.7)")
    ap.add_argument("--eager-train", action="store_true", help="train mode: launch every kernel eagerly instead of replaying a graph")
    ap.add_argument("--layers-out", default=None, help="write the per-layer roofline table (JSON) here")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the flow_net path has no CPU fallback")
    torch.cuda.set_device(local)
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries the one JSON line: NCCL's own log (its version banner when NCCL_DEBUG is set in the environment) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # the serving loop's pinned buffers should live on the GPU's NUMA node (utils/numa.py); FLOWNET_NUMA=0: leave unpinned
    from utils.numa import bind_to_gpu_numa_node
    numa = bind_to_gpu_numa_node(local) if os.environ.get("FLOWNET_NUMA", "1") != "0" else None

    import flownet_native as native
    import model.flow_architecture as arch
    import model.flow_net as flow_net

    if args.mode == "train":
        rec = run_train(args, rank, world, dist, native, arch, flow_net)
        if rec is not None:
            emit(rec)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    B = args.batch
    x_host = torch.from_numpy(car_batch(B)).pin_memory()            # uint8 [B,128,256,1]
    out_host = torch.empty((B, SHAPE[0], SHAPE[1], 2), dtype=torch.float32).pin_memory()
    arch.reset_variables(1234)
    compiled = flow_net.CompiledInference(B, SHAPE)                  # warm-up + CUDA-graph capture
    compiled(x_host.cuda())
    torch.cuda.synchronize()
    assert native.tc_status() == 0, "tcgen05 pipeline wait timed out"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (value)
    for _ in range(args.warmup):
        compiled.replay()
    barrier()
    replays0 = compiled.replays                                      # graph replays inside the three timed regions
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        compiled.replay()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)

    # ---------------- end to end through the public API with host buffers (e2e)
    # every step uploads its own pinned uint8 batch and downloads its own fp32 field; the serving loop
    # (CompiledInference.stream_host) overlaps step i's upload and step i-2's download with step i-1's compute
    out_host2 = [out_host, torch.empty_like(out_host).pin_memory()]
    xs = [x_host] * args.steps
    outs = [out_host2[i & 1] for i in range(args.steps)]
    compiled.stream_host(xs[:3], outs[:3])
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    compiled.stream_host(xs, outs)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    # the same without overlap (one stream: H2D, graph, D2H back to back), reported for reference
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(args.steps):
        out_host.copy_(compiled(x_host), non_blocking=True)
    g1.record()
    barrier()
    ms_e2e_serial = g0.elapsed_time(g1)
    timed_replays = compiled.replays - replays0 - 3                  # minus the e2e loop's three untimed warm-up steps
    launches_per_replay = compiled.launches_per_replay
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    value = world * B * args.steps / (ms * 1e-3)
    e2e = world * B * args.steps / (ms_e2e * 1e-3)

    if rank == 0:
        # ---------------- per-kernel roofline (N=1 semantics: this rank's GPU)
        rows = layer_profile(x_host.cuda())
        ridge = peaks["tc_burst"] * 1e12 / (peaks["hbm"] * 1e9)
        sm_count = torch.cuda.get_device_properties(0).multi_processor_count
        sm_hz = ((clocks or {}).get("sm_mhz") or 1965.0) * 1e6
        for r in rows:
            r["bound"] = "tensor" if r["flops"] / r["alg_bytes"] > ridge else "hbm"
            r["gbs"] = r["alg_bytes"] / (r["ms"] * 1e-3) / 1e9
            r["tflops"] = r["flops"] / (r["ms"] * 1e-3) / 1e12
            r["frac"] = (r["tflops"] / peaks["tc_burst"]) if r["bound"] == "tensor" else (r["gbs"] / peaks["hbm"])
            # MMA-rate floor: a tcgen05.mma (M=128, K=16) occupies the MMA unit 39-128 clk depending on N (DESIGN.md section 3,
            # profiles/r2_mma_bench3.log), whatever its flop count; floor = MMA-unit clocks / (SMs x SM clock)
            r["mma_floor_us"] = r.get("mma_clk", 0.0) / (sm_count * sm_hz) * 1e6
            r["floor_us"] = max(r["mma_floor_us"], r["alg_bytes"] / (peaks["hbm"] * 1e9) * 1e6, r["flops"] / (peaks["tc_burst"] * 1e12) * 1e6)
        total_ms = sum(r["ms"] for r in rows)
        top = max(rows, key=lambda r: r["ms"])
        ev = ncu_evidence() or {}
        ev_top = (ev.get("launches", {}).get("%s@b%d" % (top["name"], B)) or {})
        roofline = {
            "bound": top["bound"],
            "achieved": top["tflops"] if top["bound"] == "tensor" else top["gbs"],
            "peak": peaks["tc_burst"] if top["bound"] == "tensor" else peaks["hbm"],
            "unit": "TFLOP/s" if top["bound"] == "tensor" else "GB/s",
            "frac": top["frac"], "traffic": ev_top.get("dram_bytes"),
            "traffic_source": ev.get("source") if ev_top else None,
            # BASELINE metric, second half: tensor-pipe activity of the conv launches (ncu sm__pipe_tensor_cycles_active
            # and sm__mem_tensor_cycles_active, % of elapsed), read from the committed capture of this command
            "tensor_pipe_util": ev.get("tensor_pipe_util"),
            "kernel": top["name"], "kernel_ms": top["ms"], "share_of_step": top["ms"] / total_ms,
            "tcgen05": top["tcgen05"], "peak_source": peaks["source"] + " (burst; kernel timed alone)",
            # the launch against the larger of its three floors (HBM bytes, dense-bf16 flops, MMA-unit occupancy)
            "kernel_floor_us": top["floor_us"], "kernel_mma_floor_us": top["mma_floor_us"],
            "floor_frac": top["floor_us"] / (top["ms"] * 1e3),
            "whole_net": {
                "tflops": value / world * FLOP_PER_IMG / 1e12,
                "frac_tensor_sustained": value / world * FLOP_PER_IMG / 1e12 / peaks["tc_sustained"],
                "gbs_algorithmic": value / world * BYTES_PER_IMG / 1e9,
                "frac_hbm": value / world * BYTES_PER_IMG / 1e9 / peaks["hbm"],
                # sum over the launches of max(HBM, flop, MMA-unit) floors against the graph's step time
                "sum_floor_us": sum(r["floor_us"] for r in rows),
                "sum_mma_floor_us": sum(r["mma_floor_us"] for r in rows),
                "floor_frac": sum(r["floor_us"] for r in rows) / (ms / args.steps * 1e3),
            },
        }
        if args.layers_out:
            with open(args.layers_out, "w") as f:
                json.dump({"batch": B, "rows": rows, "sum_ms": total_ms, "graph_ms_per_step": ms / args.steps}, f,
                          indent=1)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_oracle_rate()
    # ---------------- BASELINE configs[2] beside it: the training step with its gradient all-reduce (every rank takes part)
    train = None
    if not args.no_train:
        del compiled
        torch.cuda.empty_cache()
        train = run_train(args, rank, world, dist, native, arch, flow_net, steps=args.train_steps)
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": value / PUBLISHED_IMG_S,
            "dtype": "bf16", "data": "bundled car masks tiled to 64 (fixture), random-init seed-1234 weights",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "parallelism": f"batch-sharded x{world}, no collective",
                       "l2_policy": "no flush: one step moves ~1 GB of activations per GPU, 8x the 126 MB L2",
                       "baseline_note": "vs_baseline divides by the reference README's 348 img/s (GTX 1080, batch 8, "
                                        "incl. feed/fetch); compare e2e for like with like",
                       "launch": "CUDA graph, %d kernels per step" % launches_per_replay},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(x_host.numel()) * world,
                    "d2h_bytes_per_step": int(out_host.numel() * 4) * world, "ms_per_step": ms_e2e / args.steps,
                    "how": "CompiledInference.stream_host: two alternating graphs, uploads / downloads on side streams straight into / out of their static buffers",
                    "serial_value": world * B * args.steps / (ms_e2e_serial * 1e-3),
                    # rank 0's host placement (utils/numa.py): pinned buffers allocated after binding to the GPU's NUMA node;
                    # tools/copy_ceiling.py measures what the copies alone allow at N ranks (profiles/r2_copy_ceiling*.json)
                    "host_numa": numa},
            # kernels of libflownet.so launched inside the three timed regions (value, e2e, serial e2e): counted graph
            # replays x the launches the C ABI's counter saw while the graph was captured
            "gpu_launches": launches_per_replay * timed_replays,
            "clocks": clocks,
            "roofline": roofline,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if train is not None:
            line["train"] = train
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
