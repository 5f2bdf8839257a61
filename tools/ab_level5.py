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
        return arch._level5_fused(x)
    out = {}
    for name, fn in (("layers", layers), ("fused", fused)):
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
    print("B=%d  five launches %.1f us   fused %.1f us" % (B, out["layers"], out["fused"]), flush=True)
