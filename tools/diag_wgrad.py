"""Diagnostic: tcgen05 weight-gradient kernel vs the CUDA-core one on a few shapes (both descriptor conventions)."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "steady-state-flow-with-neural-nets_b200"))
import flownet_native as native

def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))

cases = [(2, 16, 32, 8, 16, 3, 1, 1), (2, 16, 32, 16, 32, 3, 1, 0), (2, 32, 64, 8, 16, 3, 2, 1), (2, 16, 32, 64, 64, 1, 1, 1),
         (3, 8, 16, 128, 256, 3, 1, 1)]
for swap in (0, 1):
    native.lib.fn_debug_set_wgrad_swap(swap)
    for (B, H, W, C, N, k, s, pro) in cases:
        g = torch.Generator().manual_seed(3)
        x = torch.randn(B, H, W, C, generator=g).to("cuda", torch.bfloat16)
        OH, OW = -(-H // s), -(-W // s)
        dY = torch.randn(B, OH, OW, N, generator=g).to("cuda", torch.bfloat16)
        mult = 2 if pro in (1, 3) else 1
        outs = []
        for impl in (native.IMPL_TCGEN05, native.IMPL_SIMT):
            out = torch.full((k * k, C * mult, N), float("nan"), device="cuda")
            native.conv_wgrad(x, dY, out, ksize=k, stride=s, prologue=pro, impl=impl)
            torch.cuda.synchronize()
            outs.append(out)
        print("swap", swap, (B, H, W, C, N, k, s, pro), "status", native.tc_status(), "rel", rel(outs[0], outs[1]),
              "nan", int(torch.isnan(outs[0]).sum()), flush=True)
