"""A/B: data gradients of the stride-2 convs (transposed conv, Cout = Cin) at batch 256 -- persistent sub-pixel kernel
against the generic tcgen05 kernel (CUDA events, 20 launches back to back)."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "steady-state-flow-with-neural-nets_b200"))
import flownet_native as native
g = torch.Generator().manual_seed(1)
B = 256
for C, H, W in ((16, 64, 128), (32, 32, 64), (64, 16, 32)):
    x = torch.randn(B, H, W, C, generator=g).to("cuda", torch.bfloat16)
    w16 = (torch.rand(9, C, C, generator=g) * 0.03).to("cuda", torch.bfloat16)
    bias = torch.zeros(C, device="cuda")
    w_tc = native.pack_conv_weights_tc(w16, None, 3, C, 0, C, False)
    y = torch.empty((B, 2 * H, 2 * W, C), device="cuda", dtype=torch.bfloat16)
    out, ys = [], []
    for special in (0, 1):
        native.lib.fn_debug_set_use_s1p(special)
        def run():
            native.tconv_fwd(x, w16, bias, y, B=B, H=H, W=W, Cin=C, Cout=C, impl=native.IMPL_TCGEN05, w_tc=w_tc)
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
        ys.append(y.float().clone())
    native.lib.fn_debug_set_use_s1p(1)
    d = (ys[0] - ys[1]).norm() / ys[0].norm()
    gb = (x.numel() + y.numel()) * 2 / 1e9
    # the same with the concat_elu adjoint: one launch against transposed conv + fn_prologue_bwd
    z = torch.randn(B, 2 * H, 2 * W, C // 2, generator=g).to("cuda", torch.bfloat16)
    out2 = torch.zeros(B, 2 * H, 2 * W, C // 2, device="cuda", dtype=torch.bfloat16)
    def fused():
        native.tconv_fwd(x, w16, bias, out2, B=B, H=H, W=W, Cin=C, Cout=C, impl=native.IMPL_TCGEN05, w_tc=w_tc,
                         epilogue=native.EPI_PRO_BWD, res=z, adj_prologue=native.PRO_CONCAT_ELU, adj_accumulate=True)
    def two():
        native.tconv_fwd(x, w16, bias, y, B=B, H=H, W=W, Cin=C, Cout=C, impl=native.IMPL_TCGEN05, w_tc=w_tc)
        native.prologue_bwd(y, z, native.PRO_CONCAT_ELU, out=out2, accumulate=True)
    for fn in (two, fused):
        for _ in range(3):
            fn()
        torch.cuda._sleep(4_000_000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / 20 * 1000)
    print("   with the prologue adjoint: two launches %.1f us, fused %.1f us" % (out[2], out[3]))
    print("C=%d %dx%d  generic %.1f us  persistent %.1f us (%.0f GB/s)  rel diff %.1e  status %d" % (
        C, H, W, out[0], out[1], gb / (out[1] * 1e-6), d.item(), native.tc_status()), flush=True)
