#!/usr/bin/env python
"""Debug aid: per-stage clock stamps of the fused 8x16-level cluster kernel (csrc/conv_l5c.cu), thread 0 of every CTA."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "steady-state-flow-with-neural-nets_b200"))
import torch
import flownet_native as native
import model.flow_architecture as arch
B = int(os.environ.get("L5_BATCH", "64"))
arch.reset_variables(1)
x = torch.randn(B, 16, 32, 64).to("cuda", torch.bfloat16)
NCTA = 2 * B
buf = torch.zeros(NCTA * 16, dtype=torch.int64, device="cuda")
names = ["build", "S1 mma", "S1 epi+xchg", "S2 mma", "S2 epi+xchg", "S3 mma", "S3 epi+xchg", "S4 mma", "S4 epi+xchg", "S5 mma", "S5 epi"]
for _ in range(3):
    arch._level5_fused(x)
torch.cuda.synchronize()
buf.zero_()
native.lib.fn_debug_set_timing(buf.data_ptr(), NCTA * 2)
arch._level5_fused(x)
torch.cuda.synchronize()
native.lib.fn_debug_set_timing(None, 0)
st = buf.view(NCTA, 16).cpu().double()
n = len(names)
d = st[:, 1:n + 1] - st[:, 0:n]
tot = (st[:, n] - st[:, 0]).mean()
print("level 5, B=%d, mean clocks per phase over %d CTAs (total %.0f clk = %.1f us at 1.965 GHz):" % (B, NCTA, tot, tot / 1965))
for i, nm in enumerate(names):
    print("   %-12s %7.0f  (min %6.0f max %6.0f)" % (nm, d[:, i].mean(), d[:, i].min(), d[:, i].max()))
