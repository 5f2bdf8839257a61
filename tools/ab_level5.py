"""A/B: the 8x16 level as five launches vs the fused fn_level5_fwd launch (CUDA events, graph of 20 back-to-back runs)."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "steady-state-flow-with-neural-nets_b200"))
import flownet_native as native
import model.flow_architecture as arch
for B in (8, 64, 148, 256):
    arch.reset_variables(1)
    x = torch.randn(B, 16, 32, 64).to("cuda", torch.bfloat16)
    conc = arch.concat_elu
    def layers():
        t = arch.res_block(x, filter_size=128, nonlinearity=conc, stride=2, gated=True, name="resnet_5_downsample")
        t = arch.res_block(t, filter_size=128, nonlinearity=conc, gated=True, name="resnet_5_0")
        return arch.transpose_conv_layer(t, 3, 2, 64, "up_conv_1")
    def fused():
        arch.LEVEL5_SPLIT = False
        return arch._level5_fused(x)
    def fused2():
        arch.LEVEL5_SPLIT = True
        return arch._level5_fused(x)
    out = {}
    for name, fn in (("layers", layers), ("fused", fused), ("fused2", fused2)):
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
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
        out[name] = e0.elapsed_time(e1) / 100 * 1000
    assert native.tc_status() == 0
    print("B=%d  five launches %.1f us   fused %.1f us   fused, 2 CTAs per image %.1f us" % (B, out["layers"], out["fused"], out["fused2"]), flush=True)

# cold-cache timing: an L2 flush (writing 512 MB) before every run, the flush alone subtracted
if os.environ.get("AB_COLD"):
    B = 64
    arch.reset_variables(1)
    x = torch.randn(B, 16, 32, 64).to("cuda", torch.bfloat16)
    junk = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    def flush():
        junk.fill_(1)
    def timed_graph(body):
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(2):
                body()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(10):
                body()
        gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 30 * 1000
    t_flush = timed_graph(flush)
    for name, fn in (("layers", layers), ("fused", fused), ("fused2", fused2)):
        t = timed_graph(lambda: (flush(), fn()))
        print("cold B=64 %s: %.1f us (flush alone %.1f)" % (name, t - t_flush, t_flush), flush=True)
