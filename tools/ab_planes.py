"""A/B: the stride-1 concat_elu convs of the 128x256 / 64x128 / 32x64 levels with raw operands (k_conv_wg / k_conv_s1p:
TMA -> transform -> MMA -> epilogue) vs MMA-ready operands (k_conv_pl: TMA -> MMA -> epilogue), batch 64, CUDA events over
20 back-to-back launches.  Outputs as the network asks for them: conv_1 writes planes only, conv_2 raw (+ planes)."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "steady-state-flow-with-neural-nets_b200"))
import flownet_native as native
g = torch.Generator().manual_seed(1)
B = int(os.environ.get("AB_BATCH", "64"))


def to_planes(t):
    a = native.nonlinearity(t.contiguous(), native.PRO_CONCAT_ELU)
    b, h, w, c2 = a.shape
    return a.reshape(b, h, w, c2 // 8, 8).permute(0, 3, 1, 2, 4).contiguous()


def timed(run):
    for _ in range(3):
        run()
    st = native.tc_status()
    if st != 0:
        print("  !! pipeline status", st, "ne/j", os.environ.get("_TAG"), flush=True)
        native.lib.fn_debug_tc_reset() if hasattr(native.lib, "fn_debug_tc_reset") else None
        return float("nan")
    torch.cuda._sleep(4_000_000)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20 * 1000


# (H, W, C, gate, skip)
CASES = [(128, 256, 8, True, False), (128, 256, 8, False, False), (128, 256, 8, False, True),
         (64, 128, 16, True, False), (64, 128, 16, False, False), (64, 128, 16, False, True),
         (32, 64, 32, True, False), (32, 64, 32, False, False), (32, 64, 32, False, True)]
for (H, W, C, gate, skip) in CASES:
    N = 2 * C if gate else C
    x = torch.randn(B, H, W, C, generator=g).to("cuda", torch.bfloat16)
    w16 = (torch.rand(9, 2 * C, N, generator=g) * 0.03).to("cuda", torch.bfloat16)
    ws16 = (torch.rand(2 * C, N, generator=g) * 0.03).to("cuda", torch.bfloat16) if skip else None
    sk = torch.randn(B, H, W, C, generator=g).to("cuda", torch.bfloat16) if skip else None
    bias = torch.zeros(N, device="cuda")
    res = torch.randn(B, H, W, C, generator=g).to("cuda", torch.bfloat16) if gate else None
    w_tc = native.pack_conv_weights_tc(w16, ws16, 3, 2 * C, 2 * C if skip else 0, N, gate)
    y = torch.empty((B, H, W, C), device="cuda", dtype=torch.bfloat16)
    yp = torch.empty((B, 2 * C // 8, H, W, 8), device="cuda", dtype=torch.bfloat16)
    xp, skp = to_planes(x), (to_planes(sk) if skip else None)
    kw = dict(B=B, H=H, W=W, Cin=C, Cout=N, ksize=3, stride=1, prologue=1, epilogue=1 if gate else 0,
              impl=native.IMPL_TCGEN05, w_tc=w_tc, res=res, w_skip=ws16)
    out = {}
    out["raw"] = timed(lambda: native.conv_fwd(x, w16, bias, y, skip=sk, **kw))
    for ne, ni in ((0, 0), (8, 1), (8, 2), (8, 4), (16, 1), (16, 2), (16, 4)):
        for j in ((0,) if ne == 0 else (4,) if C == 8 else ((2,) if skip else (4, 2)) if C == 16 else ((1, 2) if not gate and not skip else (1,))):
            native.lib.fn_debug_set_pl_ne(ne)
            native.lib.fn_debug_set_pl_ni(ni)
            native.lib.fn_debug_set_pl_j(j)
            tag = "e%di%dj%d" % (ne, ni, j)
            os.environ["_TAG"] = tag
            if gate:
                out[tag + ">raw"] = timed(lambda: native.conv_fwd(xp, w16, bias, y, skip=skp, x_planes=True, skip_planes=skip, **kw))
                out[tag + ">raw+pl"] = timed(lambda: native.conv_fwd(xp, w16, bias, y, skip=skp, x_planes=True, skip_planes=skip, y_planes=yp, **kw))
            else:
                out[tag + ">pl"] = timed(lambda: native.conv_fwd(xp, w16, bias, None, skip=skp, x_planes=True, skip_planes=skip, y_planes=yp, **kw))
    native.lib.fn_debug_set_pl_ne(0)
    native.lib.fn_debug_set_pl_ni(0)
    native.lib.fn_debug_set_pl_j(0)
    if os.environ.get("AB_KNOCK"):
        for kn, nm in ((1, "noMMA"), (2, "noST"), (4, "noTMA"), (3, "noMMA+ST"), (5, "noMMA+TMA"), (6, "noST+TMA"), (7, "empty")):
            native.lib.fn_debug_set_pl_knock(kn)
            os.environ["_TAG"] = nm
            if gate:
                out[nm] = timed(lambda: native.conv_fwd(xp, w16, bias, y, skip=skp, x_planes=True, skip_planes=skip, **kw))
            else:
                out[nm] = timed(lambda: native.conv_fwd(xp, w16, bias, None, skip=skp, x_planes=True, skip_planes=skip, y_planes=yp, **kw))
        native.lib.fn_debug_set_pl_knock(0)
    print("B=%d %dx%d C=%d gate=%d skip=%d  " % (B, H, W, C, gate, skip) + "  ".join("%s %.1f" % kv for kv in out.items()), flush=True)
