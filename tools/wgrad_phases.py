"""(flag 128: 8x16 tiles instead of 8x32 on narrow layers)  Timing experiment: the tcgen05 wgrad kernel with individual pipeline parts disabled (results are garbage then).
flags: 2 no MMA, 4 no dY TMA, 8 no x TMA, 16 no transform."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "steady-state-flow-with-neural-nets_b200"))
import flownet_native as native

def run(B, H, W, C, N, k, s, pro, flags):
    x = torch.randn(B, H, W, C, device="cuda").to(torch.bfloat16)
    OH, OW = -(-H // s), -(-W // s)
    dY = torch.randn(B, OH, OW, N, device="cuda").to(torch.bfloat16)
    mult = 2 if pro in (1, 3) else 1
    out = torch.empty((k * k, C * mult, N), device="cuda")
    native.lib.fn_debug_set_wgrad_swap(flags)
    for _ in range(2):
        native.conv_wgrad(x, dY, out, ksize=k, stride=s, prologue=pro, impl=native.IMPL_TCGEN05)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        native.conv_wgrad(x, dY, out, ksize=k, stride=s, prologue=pro, impl=native.IMPL_TCGEN05)
    e1.record()
    torch.cuda.synchronize()
    native.lib.fn_debug_set_wgrad_swap(0)
    return e0.elapsed_time(e1) / 5 * 1000, native.tc_status()

shapes = {"L1 c8 n16": (128, 128, 256, 8, 16, 3, 1, 1), "L2 c16 n32": (128, 64, 128, 16, 32, 3, 1, 1),
          "L3 c32 n64": (256, 32, 64, 32, 64, 3, 1, 1), "L3 c32 n32": (256, 32, 64, 32, 32, 3, 1, 1),
          "S2 c8 n16": (256, 128, 256, 8, 16, 3, 2, 1), "S2 c16 n32": (256, 64, 128, 16, 32, 3, 2, 1),
          "L4 c64 n128": (256, 16, 32, 64, 128, 3, 1, 1), "L5 c128 n256": (256, 8, 16, 128, 256, 3, 1, 1)}
for name, sh in shapes.items():
    row = []
    for flags in [int(f) for f in os.environ.get('WGRAD_FLAGS', '0,128,2,4,8,16,18,12,30').split(',')]:
        t, st = run(*sh, flags)
        row.append("%d:%.0fus" % (flags, t))
    print(name, " ".join(row), flush=True)
