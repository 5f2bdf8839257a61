"""Stress / determinism check: the captured forward replayed N times on the same input must give bit-identical output every time
(a race in the fused cluster kernels -- ring slots, halo exchange, MMA-burst lock -- would show up as a mismatch or a status code)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "steady-state-flow-with-neural-nets_b200"))
import flownet_native as native
import model.flow_architecture as arch
import model.flow_net as flow_net
from oracle import flow_oracle as fo
N = int(os.environ.get("STRESS_N", "3000"))
for B in [int(v) for v in os.environ.get("STRESS_BATCHES", "64,8,37,74").split(",")]:
    arch.reset_variables(1234)
    x = torch.from_numpy(fo.synth_boundary(B, 128, 256, seed=B).astype(np.uint8)).cuda()
    run = flow_net.CompiledInference(B, (128, 256), keep_prob=1.0, in_dtype=torch.uint8)
    ref = run(x).clone()
    bad = 0
    for i in range(N):
        y = run(x)
        if i % 50 == 0 or i == N - 1:
            if not torch.equal(y, ref):
                bad += 1
    torch.cuda.synchronize()
    st = native.tc_status()
    print("B=%d: %d replays, %d mismatching checks, status %d, finite %s" % (B, N, bad, st, bool(torch.isfinite(ref).all())), flush=True)
    assert bad == 0 and st == 0
print("OK")
