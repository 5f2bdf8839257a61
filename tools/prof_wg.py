"""One stride-1 concat_elu conv launch shape for ncu: PROF_CASE = 'H,W,C,gate,skip', PROF_WG = 0/1, PROF_J16 = 1/2."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "steady-state-flow-with-neural-nets_b200"))
import flownet_native as native
g = torch.Generator().manual_seed(1)
B = int(os.environ.get("AB_BATCH", "64"))
H, W, C, gate, skip = [int(v) for v in os.environ.get("PROF_CASE", "128,256,8,1,0").split(",")]
N = 2 * C if gate else C
x = torch.randn(B, H, W, C, generator=g).to("cuda", torch.bfloat16)
w16 = (torch.rand(9, 2 * C, N, generator=g) * 0.03).to("cuda", torch.bfloat16)
ws16 = (torch.rand(2 * C, N, generator=g) * 0.03).to("cuda", torch.bfloat16) if skip else None
sk = torch.randn(B, H, W, C, generator=g).to("cuda", torch.bfloat16) if skip else None
bias = torch.zeros(N, device="cuda")
res = torch.randn(B, H, W, C, generator=g).to("cuda", torch.bfloat16) if gate else None
w_tc = native.pack_conv_weights_tc(w16, ws16, 3, 2 * C, 2 * C if skip else 0, N, bool(gate))
y = torch.empty((B, H, W, C), device="cuda", dtype=torch.bfloat16)
native.lib.fn_debug_set_wg(int(os.environ.get("PROF_WG", "2")))
native.lib.fn_debug_set_wg_j16(int(os.environ.get("PROF_J16", "1")))
for _ in range(int(os.environ.get("PROF_N", "4"))):
    native.conv_fwd(x, w16, bias, y, B=B, H=H, W=W, Cin=C, Cout=N, ksize=3, stride=1, prologue=1,
                    epilogue=1 if gate else 0, impl=native.IMPL_TCGEN05, w_tc=w_tc, res=res, skip=sk, w_skip=ws16)
torch.cuda.synchronize()
assert native.tc_status() == 0
print("ok")
