"""A/B: head (1 -> 8 channels) and tail (8 -> 2 + tanh) CUDA-core kernels, round-1 version vs the vertical-strip version."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "steady-state-flow-with-neural-nets_b200"))
import flownet_native as native
g = torch.Generator().manual_seed(1)
B, H, W = int(os.environ.get("AB_BATCH", "64")), 128, 256


def timed(run):
    for _ in range(3):
        run()
    torch.cuda._sleep(4_000_000)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20 * 1000


xh = (torch.rand(B, H, W, 1, generator=g) > 0.5).to("cuda", torch.bfloat16)
wh = (torch.rand(9, 1, 8, generator=g) - 0.5).to("cuda", torch.bfloat16)
bh = torch.rand(8, generator=g).cuda()
yh = torch.empty(B, H, W, 8, device="cuda", dtype=torch.bfloat16)
xt = torch.randn(B, H, W, 8, generator=g).to("cuda", torch.bfloat16)
wt = ((torch.rand(9, 8, 2, generator=g) - 0.5) * 0.3).to("cuda", torch.bfloat16)
bt = torch.rand(2, generator=g).cuda()
yt = torch.empty(B, H, W, 2, device="cuda", dtype=torch.float32)
res = {}
for v in (0, 1):
    native.lib.fn_debug_set_tail_v2(v)
    th = timed(lambda: native.conv_fwd(xh, wh, bh, yh, B=B, H=H, W=W, Cin=1, Cout=8, prologue=0, epilogue=0))
    tt = timed(lambda: native.conv_fwd(xt, wt, bt, yt, B=B, H=H, W=W, Cin=8, Cout=2, prologue=0, epilogue=3))
    res[v] = (yh.clone(), yt.clone())
    print("v%d: head %.1f us  tail %.1f us" % (2 if v else 1, th, tt), flush=True)
native.lib.fn_debug_set_tail_v2(1)
print("head max diff", (res[0][0].float() - res[1][0].float()).abs().max().item(), " tail max diff", (res[0][1] - res[1][1]).abs().max().item())
