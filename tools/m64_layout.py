"""Experiment: TMEM lane layout of an M = 64 accumulator (cta_group::1).  Runs the wgrad kernel with an M = 64
instruction descriptor on a Keff = 128 problem and matches every one of the 128 read-out lanes to a reference row."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "steady-state-flow-with-neural-nets_b200"))
import flownet_native as native
g = torch.Generator().manual_seed(3)
B, H, W, C, N = 2, 16, 32, 128, 16
x = torch.randn(B, H, W, C, generator=g).to("cuda", torch.bfloat16)
dY = torch.randn(B, H, W, N, generator=g).to("cuda", torch.bfloat16)
ref = torch.empty((9, C, N), device="cuda")
native.conv_wgrad(x, dY, ref, ksize=3, stride=1, prologue=0, impl=native.IMPL_SIMT)
out = torch.full((9, C, N), float("nan"), device="cuda")
native.lib.fn_debug_set_wgrad_swap(32)
native.conv_wgrad(x, dY, out, ksize=3, stride=1, prologue=0, impl=native.IMPL_TCGEN05)
native.lib.fn_debug_set_wgrad_swap(0)
torch.cuda.synchronize()
print("status", native.tc_status())
r, o = ref[4].double(), out[4].double()          # centre tap
mapping = []
for lane in range(128):
    d = ((r - o[lane][None, :]).norm(dim=1) / (r.norm(dim=1) + 1e-9))
    j = int(d.argmin())
    mapping.append((lane, j if float(d[j]) < 1e-3 else -1))
print(mapping)
