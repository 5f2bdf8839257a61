#!/usr/bin/env python
"""Debug aid: per-stage clock stamps of the fused level kernels INSIDE a whole forward at batch 64 (the kernels run slower
there than back to back in isolation: which phases pay?).  One forward per kernel with the stamp buffer registered only
around that kernel's launch."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "steady-state-flow-with-neural-nets_b200"))
import torch
import flownet_native as native
import model.flow_architecture as arch
import model.flow_net as flow_net
from utils.synthetic import load_car_boundaries
B = 64
cars = load_car_boundaries(os.path.join(ROOT, "tests", "golden", "cars_128x256.npz"))
x = torch.from_numpy(np.ascontiguousarray(cars[np.arange(B) % len(cars)])).cuda()
arch.reset_variables(1234)
for _ in range(3):
    flow_net.inference(x, 1.0)
torch.cuda.synchronize()
NCTA = 2 * B
buf = torch.zeros(NCTA * 16, dtype=torch.int64, device="cuda")
targets = {"_level4_down": ["build", "S1 mma", "S1 epi", "S2 mma", "S2 epi", "S3 mma", "S3 epi", "S4 mma", "S4 epi"],
           "_level5_fused": ["build", "S1 mma", "S1 epi", "S2 mma", "S2 epi", "S3 mma", "S3 epi", "S4 mma", "S4 epi", "S5 mma", "S5 epi"],
           "_level4_up": ["build", "S5 mma", "S5 epi", "S6 mma", "S6 epi", "S7 mma", "S7 epi"]}
for fname, names in targets.items():
    orig = getattr(arch, fname)
    def wrapped(*a, _orig=orig, **k):
        native.lib.fn_debug_set_timing(buf.data_ptr(), NCTA * 2)
        try:
            return _orig(*a, **k)
        finally:
            native.lib.fn_debug_set_timing(None, 0)
    setattr(arch, fname, wrapped)
    buf.zero_()
    torch.cuda._sleep(4_000_000)            # the host enqueues the whole forward behind this
    flow_net.inference(x, 1.0)
    torch.cuda.synchronize()
    setattr(arch, fname, orig)
    st = buf.view(NCTA, 16).cpu().double()
    n = len(names)
    d = st[:, 1:n + 1] - st[:, 0:n]
    tot = (st[:, n] - st[:, 0]).mean()
    spread = st[:, 0].max() - st[:, 0].min()
    print("%s in the forward: %.0f clk = %.1f us; CTA start spread %.0f clk; " % (fname, tot, tot / 1965, spread) +
          "  ".join("%s %.0f" % (nm, d[:, i].mean()) for i, nm in enumerate(names)))
