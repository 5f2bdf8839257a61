"""The 128x256-level data-gradient convs with the fused prologue adjoint at batch 256, alone (for ncu / timing)."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "steady-state-flow-with-neural-nets_b200"))
import flownet_native as native
g = torch.Generator().manual_seed(1)
B, H, W = 256, 128, 256
for C, N, acc in ((16, 16, False), (8, 16, True)):
    gx = torch.randn(B, H, W, C, generator=g).to("cuda", torch.bfloat16)
    w16 = (torch.rand(9, C, N, generator=g) * 0.03).to("cuda", torch.bfloat16)
    z = torch.randn(B, H, W, N // 2, generator=g).to("cuda", torch.bfloat16)
    out = torch.zeros(B, H, W, N // 2, device="cuda", dtype=torch.bfloat16)
    dA = torch.empty(B, H, W, N, device="cuda", dtype=torch.bfloat16)
    bias = torch.zeros(N, device="cuda")
    w_tc = native.pack_conv_weights_tc(w16, None, 3, C, 0, N, False)
    kw = dict(B=B, H=H, W=W, Cin=C, Cout=N, ksize=3, stride=1, prologue=0, w_tc=w_tc, impl=native.IMPL_TCGEN05)
    def fused():
        native.conv_fwd(gx, w16, bias, out, epilogue=native.EPI_PRO_BWD, res=z, adj_prologue=native.PRO_CONCAT_ELU,
                        adj_accumulate=acc, **kw)
    def plain():
        native.conv_fwd(gx, w16, bias, dA, epilogue=native.EPI_BIAS, **kw)
    def two():
        plain()
        native.prologue_bwd(dA, z, native.PRO_CONCAT_ELU, out=out, accumulate=acc)
    res = []
    for fn in (fused, plain, two):
        for _ in range(2):
            fn()
        torch.cuda._sleep(4_000_000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 10 * 1000)
    print("Cin=%d Cout=%d acc=%d: fused %.1f us, plain conv %.1f us, conv + prologue_bwd %.1f us, status %d" % (
        C, N, acc, res[0], res[1], res[2], native.tc_status()), flush=True)
