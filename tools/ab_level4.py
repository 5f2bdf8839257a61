"""A/B: the 16x32 level as seven launches vs the two fused fn_level4_fwd launches (CUDA events, graph of 20 back-to-back runs)."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "steady-state-flow-with-neural-nets_b200"))
import flownet_native as native
import model.flow_architecture as arch
arch.LEVEL4_MAX_WAVES = 100.0
for B in [int(v) for v in os.environ.get("AB_BATCHES", "8,64,74,128,256").split(",")]:
    arch.reset_variables(1)
    x3 = torch.randn(B, 32, 64, 32).to("cuda", torch.bfloat16)
    z = torch.randn(B, 16, 32, 64).to("cuda", torch.bfloat16)
    conc = arch.concat_elu
    def down_layers():
        t = arch.res_block(x3, filter_size=64, nonlinearity=conc, stride=2, gated=True, name="resnet_4_downsample")
        return arch.res_block(t, filter_size=64, nonlinearity=conc, gated=True, name="resnet_4_0")
    x4 = down_layers()
    def up_layers():
        u = arch.res_block(z, a=x4, filter_size=64, nonlinearity=conc, gated=True, name="resnet_up_1_0")
        return arch.transpose_conv_layer(u, 3, 2, 32, "up_conv_2")
    up_layers()
    out = {}
    for name, fn in (("down_layers", down_layers), ("down_fused", lambda: arch._level4_down(x3)),
                     ("up_layers", up_layers), ("up_fused", lambda: arch._level4_up(z, x4))):
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
    assert native.tc_status() == 0, native.tc_status()
    print("B=%d  down: four launches %.1f us, fused %.1f us    up: three launches %.1f us, fused %.1f us" %
          (B, out["down_layers"], out["down_fused"], out["up_layers"], out["up_fused"]), flush=True)
