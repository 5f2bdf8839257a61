#!/usr/bin/env python
"""GPU bring-up diagnostic for the tcgen05 conv kernel: runs a few fused-conv cases with both
shared-memory-descriptor conventions (LBO/SBO as documented vs swapped), each in its own process
(a faulting kernel poisons its CUDA context), and prints the error of each against the fp64 oracle
and against the CUDA-core kernel.  Usage: python tests/diag_tc.py [> gpurun_out/diag_tc.txt]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys, os
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "steady-state-flow-with-neural-nets_b200"))
sys.path.insert(0, os.path.join(%(root)r, "tests"))
import torch
import flownet_native as native
import test_gpu_parity as T
native.lib.fn_debug_set_swap_lbo_sbo(%(swap)d)
for case in T.TC_CASES:
    name, B, H, W, Cin, Cout, s, pro, epi, skip, Cres, pool = case
    try:
        y, ref = T.run_fused_conv.__wrapped__(native, native.IMPL_TCGEN05, B, H, W, Cin, Cout, s, pro, epi, skip, Cres, pool) \
            if hasattr(T.run_fused_conv, "__wrapped__") else T.run_fused_conv(native, native.IMPL_TCGEN05, B, H, W, Cin, Cout, s, pro, epi, skip, Cres, pool)
        ys, _ = T.run_fused_conv(native, native.IMPL_SIMT, B, H, W, Cin, Cout, s, pro, epi, skip, Cres, pool)
        print("swap=%(swap)d %%-24s tc_vs_oracle=%%.3e simt_vs_oracle=%%.3e tc_vs_simt=%%.3e" %% (
            name, T.rel_l2(y, ref), T.rel_l2(ys, ref), T.rel_l2(y, ys)), flush=True)
    except AssertionError as e:
        print("swap=%(swap)d %%-24s ASSERT %%s" %% (name, e), flush=True)
    except Exception as e:
        print("swap=%(swap)d %%-24s ERROR %%r" %% (name, e), flush=True)
        break
'''

for swap in (0, 1):
    code = CHILD % dict(root=ROOT, swap=swap)
    try:
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=240)
        print(r.stdout, end="")
        if r.returncode != 0:
            print(f"swap={swap}: child exit {r.returncode}\n{r.stderr[-1500:]}")
    except subprocess.TimeoutExpired as e:
        print(f"swap={swap}: TIMEOUT\n{(e.stdout or b'')[-1500:]}")
    sys.stdout.flush()
