"""Per-launch floor: one CUDA graph of 64 dependent launches of the same conv on a tiny input (each launch reads the
previous launch's output), for each kernel family; FLOWNET_PDL=0/1 in the environment selects programmatic dependent launch."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "steady-state-flow-with-neural-nets_b200"))
import flownet_native as native
g = torch.Generator().manual_seed(1)
B = int(os.environ.get("AB_BATCH", "1"))
CASES = [(128, 256, 8, True), (64, 128, 16, True), (32, 64, 32, True), (16, 32, 64, True), (8, 16, 128, True)]
for (H, W, C, gate) in CASES:
    N = 2 * C
    xs = [torch.randn(B, H, W, C, generator=g).to("cuda", torch.bfloat16) for _ in range(2)]
    w16 = (torch.rand(9, 2 * C, N, generator=g) * 0.01).to("cuda", torch.bfloat16)
    bias = torch.zeros(N, device="cuda")
    w_tc = native.pack_conv_weights_tc(w16, None, 3, 2 * C, 0, N, True)
    def run(i):
        native.conv_fwd(xs[i & 1], w16, bias, xs[(i + 1) & 1], B=B, H=H, W=W, Cin=C, Cout=N, ksize=3, stride=1, prologue=1,
                        epilogue=1, impl=native.IMPL_TCGEN05, w_tc=w_tc, res=xs[i & 1])
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(4):
            run(i)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for i in range(64):
            run(i)
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    assert native.tc_status() == 0
    print("PDL=%s B=%d %dx%d C=%d: %.2f us per launch" % (os.environ.get("FLOWNET_PDL", "0"), B, H, W, C, e0.elapsed_time(e1) / 640 * 1000), flush=True)
