#!/usr/bin/env python
"""Debug aid: per-stage clock stamps of the fused 16x32-level kernels (csrc/conv_l4.cu), thread 0 of every CTA, SM clocks."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "steady-state-flow-with-neural-nets_b200"))
import torch
import flownet_native as native
import model.flow_architecture as arch
arch.LEVEL4_MAX_WAVES = 100.0
B = int(os.environ.get("L4_BATCH", "64"))
arch.reset_variables(1)
x3 = torch.randn(B, 32, 64, 32).to("cuda", torch.bfloat16)
z = torch.randn(B, 16, 32, 64).to("cuda", torch.bfloat16)
x4 = arch._level4_down(x3)
NCTA = 2 * B
buf = torch.zeros(NCTA * 16, dtype=torch.int64, device="cuda")
names = {"down": ["build", "S1 mma", "S1 epi+xchg", "S2 mma", "S2 epi+xchg", "S3 mma", "S3 epi+xchg", "S4 mma", "S4 epi"],
         "up": ["build", "S5 mma", "S5 epi+xchg", "S6 mma", "S6 epi+xchg", "S7 mma", "S7 epi"]}
for mode, fn in (("down", lambda: arch._level4_down(x3)), ("up", lambda: arch._level4_up(z, x4))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    buf.zero_()
    native.lib.fn_debug_set_timing(buf.data_ptr(), NCTA * 2)
    fn()
    torch.cuda.synchronize()
    native.lib.fn_debug_set_timing(None, 0)
    st = buf.view(NCTA, 16).cpu().double()
    n = len(names[mode])
    d = st[:, 1:n + 1] - st[:, 0:n]
    print(mode, "B=%d, mean clocks per phase over %d CTAs (total %.0f clk = %.1f us at 1.965 GHz):" % (B, NCTA, (st[:, n] - st[:, 0]).mean(), (st[:, n] - st[:, 0]).mean() / 1965))
    for i, nm in enumerate(names[mode]):
        print("   %-12s %7.0f  (min %6.0f max %6.0f)" % (nm, d[:, i].mean(), d[:, i].min(), d[:, i].max()))
    if mode == "up":
        i = st[:, 8:16]
        print("   issuer (last gated stage, chunks 8 and 9): wait %.0f  issue 8 MMAs %.0f  commit %.0f  loop %.0f | wait %.0f  issue %.0f  commit %.0f" % (
            (i[:, 1] - i[:, 0]).mean(), (i[:, 2] - i[:, 1]).mean(), (i[:, 3] - i[:, 2]).mean(), (i[:, 4] - i[:, 3]).mean(),
            (i[:, 5] - i[:, 4]).mean(), (i[:, 6] - i[:, 5]).mean(), (i[:, 7] - i[:, 6]).mean()))
