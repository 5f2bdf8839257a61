"""A/B: the stride-1 concat_elu convs of the 128x256 / 64x128 / 32x64 levels on k_conv_s1p (role pipeline) vs k_conv_wg
(warpgroup-autonomous), batch 64, CUDA events over 20 back-to-back launches."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "steady-state-flow-with-neural-nets_b200"))
import flownet_native as native
g = torch.Generator().manual_seed(1)
B = int(os.environ.get("AB_BATCH", "64"))
# (H, W, C, gate, skip)
CASES = [(128, 256, 8, True, False), (128, 256, 8, False, True), (64, 128, 16, True, False), (64, 128, 16, False, False),
         (64, 128, 16, False, True), (32, 64, 32, False, False), (32, 64, 32, False, True)]
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
    out = {}
    for name, wgon, j16, lock in (("s1p", 0, 1, 40), ("wg", 2, 1, 40), ("wg, MMA bursts unserialised", 2, 1, 0), ("wg back-off 1 ns", 2, 1, 1),
                                  ("wg back-off 20 ns", 2, 1, 20), ("wg back-off 100 ns", 2, 1, 100), ("wg_j16=2", 2, 2, 40)):
        if j16 == 2 and C != 16:
            continue
        native.lib.fn_debug_set_wg(wgon)
        native.lib.fn_debug_set_wg_j16(j16)
        native.lib.fn_debug_set_wg_lock(lock)
        def run():
            native.conv_fwd(x, w16, bias, y, B=B, H=H, W=W, Cin=C, Cout=N, ksize=3, stride=1, prologue=1,
                            epilogue=1 if gate else 0, impl=native.IMPL_TCGEN05, w_tc=w_tc, res=res, skip=sk, w_skip=ws16)
        for _ in range(3):
            run()
        torch.cuda._sleep(4_000_000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            run()
        e1.record()
        torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / 20 * 1000
    native.lib.fn_debug_set_wg(1)
    native.lib.fn_debug_set_wg_j16(1)
    native.lib.fn_debug_set_wg_lock(40)
    assert native.tc_status() == 0
    print("B=%d %dx%d C=%d gate=%d skip=%d  " % (B, H, W, C, gate, skip) + "  ".join("%s %.1f us" % kv for kv in out.items()), flush=True)
