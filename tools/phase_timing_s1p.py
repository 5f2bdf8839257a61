#!/usr/bin/env python
"""Debug aid: phase stamps of the persistent conv_s1p kernel (third tile of each CTA), in SM clocks."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "steady-state-flow-with-neural-nets_b200"))
import torch  # noqa: E402
import flownet_native as native  # noqa: E402

CASES = [  # name, B,H,W,Cin,Cout,epi,skip
    ("L1 conv_2 gate C8", 64, 128, 256, 8, 16, 1, False),
    ("L1 conv_1+nin C8", 64, 128, 256, 8, 8, 0, True),
    ("L2 conv_2 gate C16", 64, 64, 128, 16, 32, 1, False),
    ("L3 conv_2 gate C32", 64, 32, 64, 32, 64, 1, False),
]
NCTA = 592
buf = torch.zeros(NCTA * 8, dtype=torch.int64, device="cuda")
for name, B, H, W, Cin, Cout, epi, skip in CASES:
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, H, W, Cin, generator=g).to("cuda", torch.bfloat16)
    w16 = (torch.randn(9, Cin * 2, Cout, generator=g) * 0.1).to("cuda", torch.bfloat16)
    bias = torch.randn(Cout, generator=g).cuda()
    sk = ws16 = None
    if skip:
        sk = torch.randn(B, H, W, Cin, generator=g).to("cuda", torch.bfloat16)
        ws16 = (torch.randn(Cin * 2, Cout, generator=g) * 0.1).to("cuda", torch.bfloat16)
    w_tc = native.pack_conv_weights_tc(w16, ws16, 3, Cin * 2, Cin * 2 if skip else 0, Cout, epi == 1)
    F = Cout // 2 if epi == 1 else Cout
    res = torch.randn(B, H, W, F, generator=g).to("cuda", torch.bfloat16) if epi == 1 else None
    y = torch.empty(B, H, W, F, device="cuda", dtype=torch.bfloat16)
    kw = dict(B=B, H=H, W=W, Cin=Cin, Cout=Cout, stride=1, prologue=1, epilogue=epi, impl=native.IMPL_TCGEN05,
              w_tc=w_tc, skip=sk, w_skip=ws16, res=res)
    for _ in range(3):
        native.conv_fwd(x, w16, bias, y, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    native.conv_fwd(x, w16, bias, y, **kw)
    e1.record()
    torch.cuda.synchronize()
    buf.zero_()
    native.lib.fn_debug_set_timing(buf.data_ptr(), NCTA)
    native.conv_fwd(x, w16, bias, y, **kw)
    torch.cuda.synchronize()
    native.lib.fn_debug_set_timing(None, 0)
    st = buf.view(NCTA, 8).cpu()
    live = (st[:, 0] != 0) & (st[:, 7] != 0)
    st = st[live].double()
    d = lambda a, b: (st[:, b] - st[:, a]).mean().item()
    print(f"{name:22s} kernel {e0.elapsed_time(e1)*1e3:6.1f}us ctas {int(live.sum()):4d} | FRONT wait {d(0,1):6.0f} transform {d(1,2):6.0f} "
          f"barrier {d(2,3):6.0f} issue {d(3,4):6.0f} | BACK wait {d(5,6):6.0f} epilogue {d(6,7):6.0f} | "
          f"front start -> back done {d(0,7):7.0f} clk", flush=True)
