"""A/B: levels 4 + 5 as three fused launches vs the one merged launch (CUDA events, graph of 20 back-to-back runs)."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "steady-state-flow-with-neural-nets_b200"))
import flownet_native as native
import model.flow_architecture as arch
for B in [int(v) for v in os.environ.get("AB_BATCHES", "8,64").split(",")]:
    arch.reset_variables(1)
    x3 = torch.randn(B, 32, 64, 32).to("cuda", torch.bfloat16)
    def three():
        x4 = arch._level4_down(x3)
        return arch._level4_up(arch._level5_fused(x4), x4)
    out = {}
    for name, fn in (("three", three), ("merged", lambda: arch._level45_fused(x3))):
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
    print("B=%d  levels 4 + 5: three launches %.1f us, one merged launch %.1f us" % (B, out["three"], out["merged"]), flush=True)
