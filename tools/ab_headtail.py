"""A/B of the network's first and last convolution at batch 64, 128x256: head (1 -> 8 channels) on the CUDA cores; tail
(8 -> 2 + tanh) on the CUDA cores (k_tail8_v2) against the tensor-core kernel with two taps per MMA (csrc/conv_tail.cu)."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "steady-state-flow-with-neural-nets_b200"))
import flownet_native as native
B, H, W = int(os.environ.get("AB_BATCH", "64")), 128, 256
g = torch.Generator().manual_seed(0)
x = torch.randn(B, H, W, 8, generator=g).to("cuda", torch.bfloat16)
w16 = (torch.randn(9, 8, 2, generator=g) * 0.1).to("cuda", torch.bfloat16)
bias = torch.randn(2, generator=g).cuda()
y = torch.empty(B, H, W, 2, device="cuda", dtype=torch.float32)
def tail():
    native.conv_fwd(x, w16, bias, y, B=B, H=H, W=W, Cin=8, Cout=2, ksize=3, stride=1, prologue=0, epilogue=3, impl=native.IMPL_AUTO)
def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(20):
            fn()
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 100 * 1000
out = {}
for mode in (0, 1):
    native.lib.fn_debug_set_tail_tc(mode)
    out[mode] = timed(tail)
    ref = y.clone() if mode == 0 else ref
assert native.tc_status() == 0
print("tail 8 -> 2 + tanh, B=%d: CUDA cores %.1f us, tensor cores (5 MMAs per 128 pixels) %.1f us, max |diff| %.2e" %
      (B, out[0], out[1], (y - ref).abs().max().item()))
