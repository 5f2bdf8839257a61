"""A/B of the nin skip's data gradient at batch 256: the fused CUDA-core kernel (1x1 conv + prologue adjoint, k_nin_bwd)
against the 1x1 conv on the tensor cores followed by fn_prologue_bwd (two launches)."""
import os, sys
import torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "steady-state-flow-with-neural-nets_b200"))
import flownet_native as native
import model.flow_backward as fb
B = int(os.environ.get("AB_BATCH", "256"))
PRO = native.PRO_CONCAT_ELU

def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(10):
            fn()
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 50 * 1000

for F, H, W in ((8, 128, 256), (16, 64, 128), (32, 32, 64), (64, 16, 32)):
    g = torch.Generator().manual_seed(F)
    dx = torch.randn(B, H, W, F, generator=g).to("cuda", torch.bfloat16)
    a = torch.randn(B, H, W, F, generator=g).to("cuda", torch.bfloat16)
    wn = (torch.randn(2 * F, F, generator=g) * 0.1).cuda()
    out0 = torch.empty_like(a)
    out1 = torch.empty_like(a)
    w16 = wn.t().reshape(1, F, -1).to(torch.bfloat16).contiguous()
    def fused():
        native.nin_bwd(dx, wn, a, PRO, out=out0, accumulate=False)
    def two():
        dAa = fb._conv_plain(dx, w16, wn.shape[0], 1, 1, dx.device)
        native.prologue_bwd(dAa, a, PRO, out=out1, accumulate=False)
    t0, t1 = timed(fused), timed(two)
    tc = native.last_used_tcgen05()
    rel = ((out0.float() - out1.float()).norm() / out0.float().norm()).item()
    print("F=%d %dx%d: k_nin_bwd %.1f us, 1x1 conv + prologue_bwd %.1f us (rel diff %.2e)" % (F, H, W, t0, t1, rel), flush=True)
