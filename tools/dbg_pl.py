import os, sys, torch
sys.path.insert(0, "steady-state-flow-with-neural-nets_b200"); sys.path.insert(0, ".")
import flownet_native as native
sys.path.insert(0, "tests")
import importlib
tp = importlib.import_module("test_gpu_parity")
case = [c for c in tp.S1P_CASES if c[0] == "l2_conv2_pool_pad"][0]
name, B, H, W, Cin, Cout, s, pro, epi, skip, Cres, pool = case
y, ref, yp = tp.run_fused_conv(native, native.IMPL_TCGEN05, B, H, W, Cin, Cout, s, pro, epi, skip, Cres, pool, planes_in=True, planes_out=True)
t = tp.to_planes(native, y)
d = (yp.float() != t.float()) | torch.isnan(yp.float())
print("mismatch", d.sum().item(), "of", d.numel(), "nan in yp", torch.isnan(yp.float()).sum().item())
idx = d.nonzero()[:10]
for ix in idx:
    ix = tuple(ix.tolist())
    print(ix, yp[ix].item(), t[ix].item(), y[ix[0], ix[2], ix[3], (ix[1] % 2) * 8 + ix[4]].item())
