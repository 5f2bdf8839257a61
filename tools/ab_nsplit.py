"""A/B: 8x16-level convs with and without the two-CTA channel split (CUDA events, 20 launches back to back)."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "steady-state-flow-with-neural-nets_b200"))
import flownet_native as native
g = torch.Generator().manual_seed(1)
for B in (64, 32):
    for gate in (True, False):
        C, H, W = 128, 8, 16
        N = 2 * C if gate else C
        x = torch.randn(B, H, W, C, generator=g).to("cuda", torch.bfloat16)
        w16 = (torch.rand(9, 2 * C, N, generator=g) * 0.03).to("cuda", torch.bfloat16)
        bias = torch.zeros(N, device="cuda")
        res = torch.randn(B, H, W, C, generator=g).to("cuda", torch.bfloat16) if gate else None
        w_tc = native.pack_conv_weights_tc(w16, None, 3, 2 * C, 0, N, gate)
        y = torch.empty((B, H, W, C), device="cuda", dtype=torch.bfloat16)
        out = []
        for split in (0, 1):
            native.lib.fn_debug_set_nsplit(split)
            def run():
                native.conv_fwd(x, w16, bias, y, B=B, H=H, W=W, Cin=C, Cout=N, ksize=3, stride=1, prologue=1,
                                epilogue=1 if gate else 0, impl=native.IMPL_TCGEN05, w_tc=w_tc, res=res)
            for _ in range(3):
                run()
            torch.cuda._sleep(4_000_000)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                run()
            e1.record()
            torch.cuda.synchronize()
            out.append(e0.elapsed_time(e1) / 20 * 1000)
        native.lib.fn_debug_set_nsplit(1)
        print("B=%d gate=%d  unsplit %.1f us  split %.1f us" % (B, gate, out[0], out[1]), flush=True)
