"""A/B: the 32x64 level's stride-1 convs layer by layer vs the fused fn_level3_fwd launches (CUDA events, graph of 20 runs)."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "steady-state-flow-with-neural-nets_b200"))
import flownet_native as native
import model.flow_architecture as arch
arch.LEVEL4_MAX_WAVES = 100.0
for B in [int(v) for v in os.environ.get("AB_BATCHES", "8,64").split(",")]:
    arch.reset_variables(1)
    x2 = torch.randn(B, 64, 128, 16).to("cuda", torch.bfloat16)
    z = torch.randn(B, 32, 64, 32).to("cuda", torch.bfloat16)
    conc = arch.concat_elu
    def down_layers():
        t = arch.res_block(x2, filter_size=32, nonlinearity=conc, stride=2, gated=True, name="resnet_3_downsample")
        return arch.res_block(t, filter_size=32, nonlinearity=conc, gated=True, name="resnet_3_0")
    x3 = down_layers()
    def up_layers():
        u = arch.res_block(z, a=x3, filter_size=32, nonlinearity=conc, gated=True, name="resnet_up_2_0")
        return arch.transpose_conv_layer(u, 3, 2, 16, "up_conv_3")
    up_layers()
    out = {}
    for name, fn in (("down_layers", down_layers), ("down_fused", lambda: arch._level3_down(x2)),
                     ("up_layers", up_layers), ("up_fused", lambda: arch._level3_up(z, x3))):
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
    print("B=%d  down: four launches %.1f us, two (stride-2 conv + fused) %.1f us    up: three launches %.1f us, two (conv_1 + fused) %.1f us" %
          (B, out["down_layers"], out["down_fused"], out["up_layers"], out["up_fused"]), flush=True)

# per-stage stamps of the fused kernels
B = 64
arch.reset_variables(1)
x2 = torch.randn(B, 64, 128, 16).to("cuda", torch.bfloat16)
z = torch.randn(B, 32, 64, 32).to("cuda", torch.bfloat16)
x3 = arch._level3_down(x2)
NCTA = 2 * B
buf = torch.zeros(NCTA * 16, dtype=torch.int64, device="cuda")
names = {"down": ["build", "S2 mma", "S2 epi+xchg", "S3 mma", "S3 epi+xchg", "S4 mma", "S4 epi"],
         "up": ["build", "S6 mma", "S6 epi+xchg", "S7 mma", "S7 epi"]}
for mode, fn in (("down", lambda: arch._level3_down(x2)), ("up", lambda: arch._level3_up(z, x3))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    buf.zero_()
    native.lib.fn_debug_set_timing(buf.data_ptr(), NCTA * 2)
    fn()
    torch.cuda.synchronize()
    native.lib.fn_debug_set_timing(None, 0)
    st = buf.view(NCTA, 16).cpu().double()
    n = len(names[mode])
    d = st[:, 1:n + 1] - st[:, 0:n]
    tot = (st[:, n] - st[:, 0]).mean()
    print(mode, "B=%d, fused kernel, mean clocks per phase (total %.0f clk = %.1f us):" % (B, tot, tot / 1965),
          "  ".join("%s %.0f" % (nm, d[:, i].mean()) for i, nm in enumerate(names[mode])))
