#!/usr/bin/env python
"""Debug aid: per-CTA phase durations (SM clocks) of k_conv_tc for a few layer shapes of the
batch-64 workload, from the clock64 stamps the kernel writes when a timing buffer is registered."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "steady-state-flow-with-neural-nets_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import flownet_native as native  # noqa: E402

CASES = [  # name, B,H,W,Cin,Cout,stride,pro,epi,skip
    ("L1 conv_2 gate C8 N16", 64, 128, 256, 8, 16, 1, 1, 1, False),
    ("L1 conv_1+nin C8 N8", 64, 128, 256, 8, 8, 1, 1, 0, True),
    ("L2 conv_2 gate C16 N32", 64, 64, 128, 16, 32, 1, 1, 1, False),
    ("L3 conv_2 gate C32 N64", 64, 32, 64, 32, 64, 1, 1, 1, False),
    ("L4 conv_2 gate C64 N128", 64, 16, 32, 64, 128, 1, 1, 1, False),
    ("L5 conv_2 gate C128 N256", 64, 8, 16, 128, 256, 1, 1, 1, False),
]
NCTA = 4096
buf = torch.zeros(NCTA * 8, dtype=torch.int64, device="cuda")
for name, B, H, W, Cin, Cout, s, pro, epi, skip in CASES:
    g = torch.Generator().manual_seed(0)
    mult = 2
    x = torch.randn(B, H, W, Cin, generator=g).to("cuda", torch.bfloat16)
    w16 = (torch.randn(9, Cin * mult, Cout, generator=g) * 0.1).to("cuda", torch.bfloat16)
    bias = torch.randn(Cout, generator=g).cuda()
    sk = ws16 = None
    if skip:
        sk = torch.randn(B, H, W, Cin, generator=g).to("cuda", torch.bfloat16)
        ws16 = (torch.randn(Cin * mult, Cout, generator=g) * 0.1).to("cuda", torch.bfloat16)
    w_tc = native.pack_conv_weights_tc(w16, ws16, 3, Cin * mult, Cin * mult if skip else 0, Cout, epi == 1)
    F = Cout // 2 if epi == 1 else Cout
    res = torch.randn(B, H, W, F, generator=g).to("cuda", torch.bfloat16) if epi == 1 else None
    y = torch.empty(B, H, W, F, device="cuda", dtype=torch.bfloat16)
    kw = dict(B=B, H=H, W=W, Cin=Cin, Cout=Cout, stride=s, prologue=pro, epilogue=epi, impl=native.IMPL_TCGEN05,
              w_tc=w_tc, skip=sk, w_skip=ws16, res=res)
    for _ in range(3):
        native.conv_fwd(x, w16, bias, y, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    native.conv_fwd(x, w16, bias, y, **kw)
    e1.record()
    torch.cuda.synchronize()
    t_plain = e0.elapsed_time(e1)
    buf.zero_()
    native.lib.fn_debug_set_timing(buf.data_ptr(), NCTA)
    native.conv_fwd(x, w16, bias, y, **kw)
    torch.cuda.synchronize()
    native.lib.fn_debug_set_timing(None, 0)
    st = buf.view(NCTA, 8).cpu()
    live = st[:, 0] != 0
    st = st[live].double()
    d = lambda a, b: (st[:, b] - st[:, a]).mean().item()
    print(f"{name:28s} kernel {t_plain*1e3:7.1f}us ctas {int(live.sum()):5d} | setup {d(0,1):7.0f} gather {d(1,2):7.0f} "
          f"mma-issue {d(2,3):7.0f} mma-wait(ep) {d(2,4):7.0f} epilogue {d(4,5):7.0f} teardown {d(5,6):6.0f} "
          f"total {d(0,6):7.0f} clk", flush=True)
