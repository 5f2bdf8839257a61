"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every test calls the CUDA path through
the C ABI (ctypes -> libflownet.so) and compares with the CPU oracle on the same seeded inputs.

Tolerances: whole-net output rel-L2 <= 1e-2 for the bf16-storage /
fp32-accumulate path vs the fp64 oracle; one fused layer <= 4e-3 when both sides see the same
bf16-rounded inputs and weights; fp32 elementwise kernels <= 1e-6.
"""
import os

import numpy as np
import pytest
import torch

from oracle import flow_oracle as fo

pytestmark = pytest.mark.gpu

LAYER_TOL = 4e-3
NET_TOL = 1e-2


def rel_l2(a, b):
    a = a.double().cpu()
    b = b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def native(built_lib):
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    return built_lib


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float64)


PRO_FN = {0: lambda x: x, 1: fo.concat_elu, 2: fo.elu, 3: fo.concat_relu, 4: torch.relu}


def oracle_fused_conv(x, w, b, stride, pro, epi, skip=None, ws=None, res=None, res_pool=False):
    """fp64 restatement of one fn_conv_fwd launch on bf16-rounded operands."""
    f = PRO_FN[pro]
    a = bf16_round(f(x)) if pro != 0 else x
    acc = fo.conv2d_same(a, w, stride)
    if skip is not None:
        sp = torch.zeros(acc.shape[:3] + (skip.shape[3],), dtype=torch.float64)
        sp[:, :skip.shape[1], :skip.shape[2], :] = skip
        acc = acc + bf16_round(f(sp)) @ ws
    acc = acc + b
    if epi == 0:
        return acc
    if epi == 3:
        return torch.tanh(acc)
    F_ = acc.shape[3] // 2 if epi == 1 else acc.shape[3]
    out = acc[..., :F_] * torch.sigmoid(acc[..., F_:]) if epi == 1 else acc
    if res is not None:
        r = fo.avg_pool_2x2_same(res) if res_pool else res
        if r.shape[3] != F_:
            r = torch.nn.functional.pad(r, (F_ - r.shape[3], 0))
        out = out + r
    return out


def to_planes(native, t):
    """planes form (include/flownet.h) of a raw bf16 NHWC CUDA tensor, built from the standalone concat_elu kernel:
    [B,H,W,2C] -> [B, 2C/8, H, W, 8]"""
    a = native.nonlinearity(t.contiguous(), native.PRO_CONCAT_ELU)
    B, H, W, C2 = a.shape
    return a.reshape(B, H, W, C2 // 8, 8).permute(0, 3, 1, 2, 4).contiguous()


def planes_match(native, yp, y):
    """yp == concat_elu(y) in planes form.  The kernels' epilogues and the standalone concat_elu kernel evaluate
    exp(-|v|) - 1 with different instruction schedules, so single elements near v = 0 (cancellation: the fp32 error of
    exp, ~1e-7, against |v|) may differ by a bf16 ulp or two; everything else is bit-identical."""
    t = to_planes(native, y).float()
    d = (yp.float() - t).abs()
    assert not torch.isnan(yp.float()).any()
    bad = d > 2.0 ** -6 * t.abs() + 2e-7
    frac = (d != 0).float().mean().item()
    if bad.any() or frac >= 1e-3:
        ix = tuple(bad.nonzero()[0].tolist()) if bad.any() else tuple((d != 0).nonzero()[0].tolist())
        print("planes mismatch: %d beyond one ulp, %.2e differ; first at %s: %g vs %g" % (
            int(bad.sum()), frac, ix, yp[ix].item(), t[ix].item()))
        return False
    return True


def run_fused_conv(native, impl, B, H, W, Cin, Cout, stride, pro, epi, with_skip=False, Cres=None, res_pool=False,
                   seed=0, ksize=3, planes_in=False, planes_out=False, raw_out=True):
    g = torch.Generator().manual_seed(seed)
    mult = 2 if pro in (1, 3) else 1
    x = bf16_round(torch.randn(B, H, W, Cin, generator=g, dtype=torch.float64))
    lim = (6.0 / (ksize * ksize * (Cin * mult + Cout))) ** 0.5
    w = bf16_round((torch.rand(ksize, ksize, Cin * mult, Cout, generator=g, dtype=torch.float64) * 2 - 1) * lim * 2)
    b = (torch.rand(Cout, generator=g, dtype=torch.float64) * 2 - 1).float().double()
    OH, OW = -(-H // stride), -(-W // stride)
    skip = ws = res = None
    if with_skip:
        Cs = Cin
        skip = bf16_round(torch.randn(B, OH, OW, Cs, generator=g, dtype=torch.float64))
        ws = bf16_round((torch.rand(Cs * mult, Cout, generator=g, dtype=torch.float64) * 2 - 1) * 0.3)
    F_ = Cout // 2 if epi == 1 else Cout
    if epi in (1, 2):
        Cres = F_ if Cres is None else Cres
        RH, RW = (2 * OH, 2 * OW) if res_pool else (OH, OW)
        res = bf16_round(torch.randn(B, RH, RW, Cres, generator=g, dtype=torch.float64))
    ref = oracle_fused_conv(x, w, b, stride, pro, epi, skip, ws, res, res_pool)

    dev = "cuda"
    xd = x.to(dev, torch.bfloat16)
    w16 = w.reshape(ksize * ksize, Cin * mult, Cout).to(dev, torch.bfloat16).contiguous()
    ws16 = ws.to(dev, torch.bfloat16).contiguous() if ws is not None else None
    bd = b.to(dev, torch.float32)
    w_tc = None
    if (Cin * mult) % 16 == 0 or (Cin * mult == 8 and not with_skip):
        w_tc = native.pack_conv_weights_tc(w16, ws16, ksize, Cin * mult, (skip.shape[3] * mult) if with_skip else 0,
                                           Cout, epi == 1)
    y = torch.empty((B, OH, OW, F_), device=dev, dtype=torch.float32 if epi == 3 else torch.bfloat16) if raw_out else None
    yp = torch.full((B, 2 * F_ // 8, OH, OW, 8), float("nan"), device=dev, dtype=torch.bfloat16) if planes_out else None
    sd = skip.to(dev, torch.bfloat16) if with_skip else None
    if planes_in:        # the operands in MMA-ready form (csrc/conv_pl.cu)
        xd = to_planes(native, xd)
        sd = to_planes(native, sd) if sd is not None else None
    native.conv_fwd(xd, w16, bd, y, B=B, H=H, W=W, Cin=Cin, Cout=Cout, ksize=ksize, stride=stride, prologue=pro,
                    epilogue=epi, impl=impl, w_tc=w_tc, skip=sd, w_skip=ws16,
                    res=res.to(dev, torch.bfloat16) if res is not None else None, res_pool=res_pool,
                    x_planes=planes_in, skip_planes=planes_in and with_skip, y_planes=yp)
    torch.cuda.synchronize()
    assert native.tc_status() == 0
    if planes_out:
        return y, ref, yp
    return y, ref


# (name, B,H,W,Cin,Cout,stride,pro,epi,skip,Cres,pool) -- the layer shapes of SURVEY App. A, scaled down
SIMT_CASES = [
    ("head_c1", 2, 16, 32, 1, 8, 1, 0, 0, False, None, False),
    ("head_c1_ragged", 3, 9, 13, 1, 8, 1, 0, 0, False, None, False),
    ("head_c1_wide", 2, 24, 300, 1, 8, 1, 0, 0, False, None, False),
    ("head_c1_n16", 2, 24, 140, 1, 16, 1, 0, 0, False, None, False),
    ("tail_tanh_c16", 2, 20, 150, 16, 2, 1, 0, 3, False, None, False),
    ("conv1_ce", 2, 16, 32, 8, 8, 1, 1, 0, False, None, False),
    ("conv2_gate", 2, 16, 32, 8, 16, 1, 1, 1, False, None, False),
    ("conv2_gate_head_res", 2, 16, 32, 8, 16, 1, 1, 1, False, 1, False),
    ("down_conv1_s2", 2, 16, 32, 8, 16, 2, 1, 0, False, None, False),
    ("down_conv1_s2_odd", 1, 15, 21, 8, 16, 2, 1, 0, False, None, False),
    ("down_conv2_gate_pool_pad", 2, 8, 16, 16, 32, 1, 1, 1, False, 8, True),
    ("up_conv1_skip", 2, 16, 32, 16, 16, 1, 1, 0, True, None, False),
    ("deep_gate", 1, 8, 16, 64, 128, 1, 1, 1, False, None, False),
    ("tail_tanh", 2, 16, 32, 8, 2, 1, 0, 3, False, None, False),
    ("elu_res_ungated", 1, 16, 16, 16, 16, 1, 2, 2, False, None, False),
    ("crelu_gate", 1, 16, 16, 8, 16, 1, 3, 1, False, None, False),
    ("relu_bias", 1, 16, 16, 16, 8, 1, 4, 0, False, None, False),
    ("nin_1x1", 1, 16, 16, 16, 8, 1, 0, 0, False, None, False),
]


@pytest.mark.parametrize("case", SIMT_CASES, ids=[c[0] for c in SIMT_CASES])
def test_fused_conv_simt_vs_oracle(native, case):
    name, B, H, W, Cin, Cout, s, pro, epi, skip, Cres, pool = case
    k = 1 if name == "nin_1x1" else 3
    y, ref = run_fused_conv(native, native.IMPL_SIMT, B, H, W, Cin, Cout, s, pro, epi, skip, Cres, pool, ksize=k)
    assert rel_l2(y, ref) < LAYER_TOL


TC_CASES = [
    # name, B,H,W,Cin,Cout,stride,pro,epi,skip,Cres,pool
    ("conv1_ce_k16_n16", 2, 16, 32, 8, 16, 1, 1, 0, False, None, False),
    ("conv1_ce_n8_padded", 2, 16, 32, 8, 8, 1, 1, 0, False, None, False),
    ("conv2_gate_f8", 2, 16, 32, 8, 16, 1, 1, 1, False, None, False),
    ("conv2_gate_head_res", 2, 16, 32, 8, 16, 1, 1, 1, False, 1, False),
    ("conv2_gate_pool_pad", 2, 16, 16, 16, 32, 1, 1, 1, False, 8, True),
    ("conv1_skip_nin", 2, 16, 32, 16, 16, 1, 1, 0, True, None, False),
    ("wide_J4", 1, 32, 64, 8, 16, 1, 1, 1, False, None, False),
    ("ragged_edges", 1, 20, 28, 8, 16, 1, 1, 0, False, None, False),
    ("transposed_8x16", 3, 8, 16, 32, 64, 1, 1, 1, False, None, False),
    ("deep_c64_n128_ring", 2, 16, 16, 64, 128, 1, 1, 1, False, None, False),
    ("deep_c128_n256_ring", 1, 8, 16, 128, 256, 1, 1, 1, False, None, False),
    ("deep_c128_n128_skip", 1, 16, 16, 64, 64, 1, 1, 0, True, None, False),
    ("elu_k16", 1, 16, 16, 16, 16, 1, 2, 2, False, None, False),
    ("crelu_gate", 1, 16, 16, 8, 16, 1, 3, 1, False, None, False),
    ("raw_k16", 1, 16, 16, 16, 16, 1, 0, 0, False, None, False),
    ("s2_c8", 2, 32, 32, 8, 16, 2, 1, 0, False, None, False),
    ("s2_c16_J", 1, 64, 128, 16, 32, 2, 1, 0, False, None, False),
    ("s2_odd_dims", 1, 31, 33, 8, 16, 2, 1, 0, False, None, False),
    ("s2_transposed_out8x16", 2, 16, 32, 32, 64, 2, 1, 0, False, None, False),
    ("s2_c64", 1, 32, 32, 64, 64, 2, 1, 0, False, None, False),
    # 8-channel gradients of the backward pass: one k-step with a zero second half
    ("raw_k8", 2, 32, 64, 8, 16, 1, 0, 0, False, None, False),
    ("raw_k8_n8", 2, 20, 28, 8, 8, 1, 0, 0, False, None, False),
    ("raw_k8_s2", 2, 32, 64, 8, 16, 2, 0, 0, False, None, False),
    # data gradients of the transposed convs: stride 2 on raw dY (conv_s2t.cu MODE_S2N)
    ("raw_s2_c16", 2, 32, 64, 16, 32, 2, 0, 0, False, None, False),
    ("raw_s2_c32", 3, 32, 32, 32, 64, 2, 0, 0, False, None, False),
    ("raw_s2_c8_ragged_even", 2, 36, 44, 8, 16, 2, 0, 0, False, None, False),
    ("raw_s2_c16_many_tiles", 40, 64, 64, 16, 32, 2, 0, 0, False, None, False),
    # last_conv + tanh on the tensor cores (csrc/conv_tail.cu: two taps per MMA): whole tiles, ragged edges, many tiles
    ("tail_tanh_tc", 2, 32, 64, 8, 2, 1, 0, 3, False, None, False),
    ("tail_tanh_tc_ragged", 3, 45, 70, 8, 2, 1, 0, 3, False, None, False),
    ("tail_tanh_tc_many_tiles", 20, 128, 256, 8, 2, 1, 0, 3, False, None, False),
]


@pytest.mark.parametrize("case", TC_CASES, ids=[c[0] for c in TC_CASES])
def test_fused_conv_tcgen05_vs_oracle(native, case):
    name, B, H, W, Cin, Cout, s, pro, epi, skip, Cres, pool = case
    y, ref = run_fused_conv(native, native.IMPL_TCGEN05, B, H, W, Cin, Cout, s, pro, epi, skip, Cres, pool)
    assert native.last_used_tcgen05()
    assert rel_l2(y, ref) < LAYER_TOL


S1P_CASES = [
    # shapes the persistent TMA kernel (conv_s1p.cu) takes: concat_elu, C in {8,16}, Cout in {C, 2C gated}
    ("l1_conv2_gate", 3, 32, 64, 8, 16, 1, 1, 1, False, None, False),
    ("l1_conv2_head_res", 2, 32, 64, 8, 16, 1, 1, 1, False, 1, False),
    ("l1_conv1_plain", 2, 32, 64, 8, 8, 1, 1, 0, False, None, False),
    ("l1_conv1_skip", 2, 32, 64, 8, 8, 1, 1, 0, True, None, False),
    ("l2_conv2_gate", 2, 32, 64, 16, 32, 1, 1, 1, False, None, False),
    ("l2_conv2_pool_pad", 2, 32, 32, 16, 32, 1, 1, 1, False, 8, True),
    ("l2_conv1_skip", 2, 32, 64, 16, 16, 1, 1, 0, True, None, False),
    ("ragged_gate", 1, 20, 28, 8, 16, 1, 1, 1, False, None, False),
    ("ragged_skip", 2, 23, 41, 16, 16, 1, 1, 0, True, None, False),
    ("many_tiles_persistent", 40, 64, 64, 8, 16, 1, 1, 1, False, None, False),
    ("w16_J2", 2, 16, 16, 8, 16, 1, 1, 1, False, None, False),
    ("l3_conv2_gate_c32", 4, 32, 64, 32, 64, 1, 1, 1, False, None, False),
    ("l3_conv1_c32", 4, 32, 64, 32, 32, 1, 1, 0, False, None, False),
    ("l3_conv1_skip_c32", 3, 32, 64, 32, 32, 1, 1, 0, True, None, False),
    ("l3_pool_pad_c32", 2, 32, 32, 32, 64, 1, 1, 1, False, 16, True),
    # deep levels: streamed weights (conv_s1d.cu)
    ("l4_conv2_gate_c64", 4, 16, 32, 64, 128, 1, 1, 1, False, None, False),
    ("l4_conv1_c64", 4, 16, 32, 64, 64, 1, 1, 0, False, None, False),
    ("l4_conv1_skip_c64", 3, 16, 32, 64, 64, 1, 1, 0, True, None, False),
    ("l4_pool_pad_c64", 2, 16, 32, 64, 128, 1, 1, 1, False, 32, True),
    ("l4_two_tiles_per_cta", 160, 16, 16, 64, 128, 1, 1, 1, False, None, False),
    ("l5_conv2_gate_c128_tr", 5, 8, 16, 128, 256, 1, 1, 1, False, None, False),
    ("l5_conv1_c128_tr", 5, 8, 16, 128, 128, 1, 1, 0, False, None, False),
    ("l5_pool_pad_c128_tr", 3, 8, 16, 128, 256, 1, 1, 1, False, 64, True),
    ("l5_two_tiles_per_cta_tr", 150, 8, 16, 128, 128, 1, 1, 0, False, None, False),
    ("c128_16x16_gate", 2, 16, 16, 128, 256, 1, 1, 1, False, None, False),
    ("c64_8x16_tr_skip", 2, 8, 16, 64, 64, 1, 1, 0, True, None, False),
    ("single_tile", 1, 16, 8, 8, 16, 1, 1, 1, False, None, False),
    ("three_tiles_per_cta", 148 * 3, 16, 32, 8, 16, 1, 1, 1, False, None, False),
    # stride-2 conv_1 of the down-sampling blocks: parity sub-plane kernel (conv_s2t.cu)
    ("s2_c8", 3, 64, 64, 8, 16, 2, 1, 0, False, None, False),
    ("s2_c16", 2, 32, 64, 16, 32, 2, 1, 0, False, None, False),
    ("s2_c32", 2, 32, 32, 32, 64, 2, 1, 0, False, None, False),
    ("s2_ragged_even", 2, 36, 44, 8, 16, 2, 1, 0, False, None, False),
    ("s2_ragged_even_c16", 1, 38, 50, 16, 32, 2, 1, 0, False, None, False),
    ("s2_many_tiles", 40, 64, 64, 8, 16, 2, 1, 0, False, None, False),
    ("s2_many_tiles_c32", 80, 32, 32, 32, 64, 2, 1, 0, False, None, False),
    # data-gradient convs of the backward pass: no prologue, Cout = C or 2C (conv_s2t.cu MODE_S1N)
    ("s1n_c8_n8", 2, 32, 64, 8, 8, 1, 0, 0, False, None, False),
    ("s1n_c8_n16", 40, 32, 64, 8, 16, 1, 0, 0, False, None, False),
    ("s1n_c16_n16", 2, 32, 64, 16, 16, 1, 0, 0, False, None, False),
    ("s1n_c16_n32_ragged", 2, 23, 41, 16, 32, 1, 0, 0, False, None, False),
    ("s1n_c32_n32", 3, 32, 64, 32, 32, 1, 0, 0, False, None, False),
    ("s1n_c32_n64", 40, 32, 32, 32, 64, 1, 0, 0, False, None, False),
    ("s1n_c64_n64", 3, 32, 64, 64, 64, 1, 0, 0, False, None, False),
]


@pytest.mark.parametrize("case", S1P_CASES, ids=[c[0] for c in S1P_CASES])
def test_persistent_tma_kernel_vs_generic_and_oracle(native, case):
    name, B, H, W, Cin, Cout, s, pro, epi, skip, Cres, pool = case
    try:
        native.lib.fn_debug_set_use_s1p(1)
        y, ref = run_fused_conv(native, native.IMPL_TCGEN05, B, H, W, Cin, Cout, s, pro, epi, skip, Cres, pool)
        native.lib.fn_debug_set_use_s1p(0)
        yg, _ = run_fused_conv(native, native.IMPL_TCGEN05, B, H, W, Cin, Cout, s, pro, epi, skip, Cres, pool)
    finally:
        native.lib.fn_debug_set_use_s1p(1)
    assert rel_l2(y, ref) < LAYER_TOL
    assert rel_l2(y, yg) < 5e-4          # same MMA order; the gate uses tanh.approx here, ex2+rcp there


WG_CASES = [c for c in S1P_CASES if c[6] == 1 and c[7] == 1 and c[4] in (8, 16, 32) and not (c[4] == 32 and c[8] == 1)]


@pytest.mark.parametrize("j16", [1, 2])
@pytest.mark.parametrize("case", WG_CASES, ids=[c[0] for c in WG_CASES])
def test_warpgroup_kernel_matches_the_role_pipelined_kernel(native, case, j16):
    """conv_wg.cu re-orchestrates k_conv_s1p (same operands, same MMA order per accumulator): the outputs agree up to
    the fused multiply-add the compiler forms in the warpgroup kernel's shortcut add, for both tile widths of the
    16-channel level."""
    name, B, H, W, Cin, Cout, s, pro, epi, skip, Cres, pool = case
    if j16 == 2 and Cin != 16:
        pytest.skip("tile-width switch only exists on the 16-channel level")
    try:
        native.lib.fn_debug_set_wg(2)          # every shape the kernel has, not only the ones it is the default for
        native.lib.fn_debug_set_wg_j16(j16)
        y, ref = run_fused_conv(native, native.IMPL_TCGEN05, B, H, W, Cin, Cout, s, pro, epi, skip, Cres, pool)
        native.lib.fn_debug_set_wg(0)
        y0, _ = run_fused_conv(native, native.IMPL_TCGEN05, B, H, W, Cin, Cout, s, pro, epi, skip, Cres, pool)
    finally:
        native.lib.fn_debug_set_wg(1)
        native.lib.fn_debug_set_wg_j16(1)
    assert rel_l2(y, ref) < LAYER_TOL
    assert rel_l2(y, y0) < 5e-4


PL_CASES = [c for c in S1P_CASES if c[6] == 1 and c[7] == 1 and c[4] in (8, 16, 32)] + [
    ("pl_wide_partial_tiles", 3, 48, 200, 8, 16, 1, 1, 1, False, None, False),
    ("pl_c16_skip_partial", 2, 40, 72, 16, 16, 1, 1, 0, True, None, False),
    ("pl_c32_many", 70, 32, 64, 32, 64, 1, 1, 1, False, None, False),
    ("pl_c8_many_skip", 33, 64, 128, 8, 8, 1, 1, 0, True, None, False),
]


@pytest.mark.parametrize("case", PL_CASES, ids=[c[0] for c in PL_CASES])
def test_planes_kernel_is_bit_identical_to_the_raw_input_kernels(native, case):
    """csrc/conv_pl.cu: the same convolution with its operands given in planes form (concat_elu applied by the producer)
    -- same weights image, same MMA order, same epilogue arithmetic as k_conv_wg / k_conv_s1p on the raw tensor, so the raw
    output is bit-identical; the planes output is concat_elu of that bf16 output."""
    name, B, H, W, Cin, Cout, s, pro, epi, skip, Cres, pool = case
    y, ref, yp = run_fused_conv(native, native.IMPL_TCGEN05, B, H, W, Cin, Cout, s, pro, epi, skip, Cres, pool,
                                planes_in=True, planes_out=True)
    y0, _ = run_fused_conv(native, native.IMPL_TCGEN05, B, H, W, Cin, Cout, s, pro, epi, skip, Cres, pool)
    assert rel_l2(y, ref) < LAYER_TOL
    assert rel_l2(y, y0) < 5e-4
    assert planes_match(native, yp, y)
    # planes only (what conv_1 writes in the network)
    _, _, yp2 = run_fused_conv(native, native.IMPL_TCGEN05, B, H, W, Cin, Cout, s, pro, epi, skip, Cres, pool,
                               planes_in=True, planes_out=True, raw_out=False)
    assert torch.equal(yp2, yp)


@pytest.mark.parametrize("case", [c for c in S1P_CASES if c[0].startswith("s2_")] + [SIMT_CASES[0], SIMT_CASES[2], SIMT_CASES[3]],
                         ids=lambda c: c[0])
def test_planes_output_of_the_stride2_and_head_kernels(native, case):
    name, B, H, W, Cin, Cout, s, pro, epi, skip, Cres, pool = case
    impl = native.IMPL_AUTO if Cin == 1 else native.IMPL_TCGEN05
    y, ref, yp = run_fused_conv(native, impl, B, H, W, Cin, Cout, s, pro, epi, skip, Cres, pool, planes_out=True)
    assert rel_l2(y, ref) < LAYER_TOL
    assert planes_match(native, yp, y)
    _, _, yp2 = run_fused_conv(native, impl, B, H, W, Cin, Cout, s, pro, epi, skip, Cres, pool, planes_out=True, raw_out=False)
    assert torch.equal(yp2, yp)


def test_planes_operands_are_refused_where_no_kernel_takes_them(native):
    with pytest.raises(ValueError):      # 64 channels: no planes kernel
        run_fused_conv(native, native.IMPL_TCGEN05, 1, 16, 32, 64, 128, 1, 1, 1, planes_in=True)
    with pytest.raises(ValueError):      # planes output from a stride-1 raw-input conv
        run_fused_conv(native, native.IMPL_TCGEN05, 1, 32, 64, 8, 16, 1, 1, 1, planes_out=True)
    with pytest.raises(ValueError):      # CUDA-core implementation asked for
        run_fused_conv(native, native.IMPL_SIMT, 1, 32, 64, 8, 16, 1, 1, 1, planes_in=True)


def test_tcgen05_refuses_shapes_it_has_no_kernel_for(native):
    with pytest.raises(ValueError):
        run_fused_conv(native, native.IMPL_TCGEN05, 1, 16, 16, 1, 8, 1, 0, 0)     # K = 9: CUDA-core layer


@pytest.mark.parametrize("impl", ["simt", "tc"])
@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 8, 16, 16, 8), (1, 4, 8, 128, 64), (2, 5, 7, 8, 8), (1, 16, 32, 32, 16),
                                            (2, 32, 64, 16, 8), (1, 20, 12, 64, 32),
                                            (2, 32, 64, 16, 16), (1, 16, 32, 32, 32), (1, 20, 12, 64, 64)])
def test_transposed_conv_vs_oracle(native, impl, B, H, W, Cin, Cout):
    if impl == "tc" and Cin % 16 != 0:
        pytest.skip("K per tap not a multiple of 16: CUDA-core layer")
    g = torch.Generator().manual_seed(1)
    x = bf16_round(torch.randn(B, H, W, Cin, generator=g, dtype=torch.float64))
    w = bf16_round((torch.rand(3, 3, Cout, Cin, generator=g, dtype=torch.float64) * 2 - 1) * 0.2)
    b = torch.rand(Cout, generator=g, dtype=torch.float64).float().double()
    ref = fo.conv2d_transpose_same(x, w, 2) + b
    w16 = w.permute(0, 1, 3, 2).reshape(9, Cin, Cout).to("cuda", torch.bfloat16).contiguous()
    w_tc = native.pack_conv_weights_tc(w16, None, 3, Cin, 0, Cout, False) if Cin % 16 == 0 else None
    y = torch.empty((B, 2 * H, 2 * W, Cout), device="cuda", dtype=torch.bfloat16)
    native.tconv_fwd(x.to("cuda", torch.bfloat16), w16, b.to("cuda", torch.float32), y, B=B, H=H, W=W, Cin=Cin,
                     Cout=Cout, impl=native.IMPL_TCGEN05 if impl == "tc" else native.IMPL_SIMT, w_tc=w_tc)
    torch.cuda.synchronize()
    assert native.tc_status() == 0
    assert native.last_used_tcgen05() == (impl == "tc")
    assert rel_l2(y, ref) < LAYER_TOL


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(3, 64, 128, 16, 8), (2, 32, 64, 32, 16), (2, 16, 32, 64, 32), (1, 20, 12, 64, 32),
                                            (40, 32, 32, 16, 8), (70, 16, 16, 32, 16), (2, 19, 27, 16, 8),
                                            # Cout = Cin: data gradients of the stride-2 convs
                                            (3, 64, 128, 16, 16), (2, 32, 64, 32, 32), (2, 16, 32, 64, 64), (40, 32, 32, 16, 16),
                                            (70, 16, 16, 32, 32), (2, 19, 27, 16, 16), (150, 16, 16, 64, 64)])
def test_transposed_conv_persistent_vs_generic(native, B, H, W, Cin, Cout):
    """conv_s2t.cu's sub-pixel kernel against the generic tcgen05 kernel (same MMA order: bit-identical)
    and the oracle."""
    g = torch.Generator().manual_seed(2)
    x = bf16_round(torch.randn(B, H, W, Cin, generator=g, dtype=torch.float64))
    w = bf16_round((torch.rand(3, 3, Cout, Cin, generator=g, dtype=torch.float64) * 2 - 1) * 0.2)
    b = torch.rand(Cout, generator=g, dtype=torch.float64).float().double()
    w16 = w.permute(0, 1, 3, 2).reshape(9, Cin, Cout).to("cuda", torch.bfloat16).contiguous()
    w_tc = native.pack_conv_weights_tc(w16, None, 3, Cin, 0, Cout, False)
    ys = []
    try:
        for use in (1, 0):
            native.lib.fn_debug_set_use_s1p(use)
            y = torch.full((B, 2 * H, 2 * W, Cout), float("nan"), device="cuda", dtype=torch.bfloat16)
            # the persistent kernel also writes the planes form of its output (consumed by conv_pl.cu)
            yp = torch.full((B, Cout // 4, 2 * H, 2 * W, 8), float("nan"), device="cuda", dtype=torch.bfloat16) if use else None
            native.tconv_fwd(x.to("cuda", torch.bfloat16), w16, b.to("cuda", torch.float32), y, B=B, H=H, W=W, Cin=Cin,
                             Cout=Cout, impl=native.IMPL_TCGEN05, w_tc=w_tc, y_planes=yp)
            torch.cuda.synchronize()
            assert native.tc_status() == 0
            if yp is not None:
                assert planes_match(native, yp, y)
            ys.append(y)
    finally:
        native.lib.fn_debug_set_use_s1p(1)
    if B * H * W <= 8192:
        assert rel_l2(ys[0], fo.conv2d_transpose_same(x, w, 2) + b) < LAYER_TOL
    assert torch.equal(ys[0], ys[1])


@pytest.mark.parametrize("B,H,W,C", [(3, 128, 256, 8), (2, 64, 128, 16), (2, 32, 64, 32), (2, 36, 44, 8), (40, 64, 64, 16),
                                     (150, 32, 32, 32)])
def test_raw_stride2_persistent_vs_generic(native, B, H, W, C):
    """conv_s2t.cu MODE_S2N (data gradient of a transposed conv: stride 2, no prologue, Cout = 2 Cin) against the generic
    tcgen05 kernel: same MMA order, bit-identical."""
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, H, W, C, generator=g).to("cuda", torch.bfloat16)
    w16 = ((torch.rand(9, C, 2 * C, generator=g) * 2 - 1) * 0.2).to("cuda", torch.bfloat16)
    b = torch.rand(2 * C, generator=g).to("cuda")
    w_tc = native.pack_conv_weights_tc(w16, None, 3, C, 0, 2 * C, False)
    ys = []
    try:
        for use in (1, 0):
            native.lib.fn_debug_set_use_s1p(use)
            y = torch.full((B, H // 2, W // 2, 2 * C), float("nan"), device="cuda", dtype=torch.bfloat16)
            native.conv_fwd(x, w16, b, y, B=B, H=H, W=W, Cin=C, Cout=2 * C, ksize=3, stride=2, prologue=native.PRO_NONE,
                            epilogue=native.EPI_BIAS, impl=native.IMPL_TCGEN05, w_tc=w_tc)
            torch.cuda.synchronize()
            assert native.tc_status() == 0
            ys.append(y)
    finally:
        native.lib.fn_debug_set_use_s1p(1)
    assert torch.equal(ys[0], ys[1])


# ------------------------------------------------------------------------------- whole network
def _net(native, x_np, impl, **flags):
    import model.flow_architecture as arch
    import model.flow_net as flow_net
    arch.reset_variables(1234)
    arch.set_impl(impl)
    old = {k: getattr(flow_net.FLAGS, k) for k in flags}
    try:
        for k, v in flags.items():
            setattr(flow_net.FLAGS, k, v)
        y = flow_net.inference(torch.from_numpy(x_np).cuda(), 1.0)
        torch.cuda.synchronize()
        assert native.tc_status() == 0
        return y
    finally:
        for k, v in old.items():
            setattr(flow_net.FLAGS, k, v)
        arch.set_impl("auto")


@pytest.mark.parametrize("impl", ["simt", "auto"])
def test_net_vs_golden_tiny(native, golden_dir, impl):
    z = np.load(os.path.join(golden_dir, "golden_tiny.npz"))
    y = _net(native, z["x"], impl)
    assert y.dtype == torch.float32 and tuple(y.shape) == z["y"].shape
    assert rel_l2(y, torch.from_numpy(z["y"])) < NET_TOL
    y2 = _net(native, z["x2"], impl)
    assert rel_l2(y2, torch.from_numpy(z["y2"])) < NET_TOL


@pytest.mark.parametrize("impl", ["simt", "auto"])
def test_net_vs_golden_cars(native, golden_dir, impl):
    from utils.synthetic import load_car_boundaries
    z = np.load(os.path.join(golden_dir, "golden_cars.npz"))
    cars = load_car_boundaries(os.path.join(golden_dir, "cars_128x256.npz"))
    y = _net(native, cars[list(z["idx"])], impl)
    e = rel_l2(y, torch.from_numpy(z["y"]))
    assert e < NET_TOL, e
    assert (y.cpu() - torch.from_numpy(z["y"])).abs().max().item() < 6e-2


def test_net_variables_match_reference_inventory(native):
    import model.flow_architecture as arch
    _net(native, fo.synth_boundary(1, 16, 32).astype(np.uint8), "auto")
    vs = fo.VariableStore(1234, torch.float32)
    fo.inference(vs, torch.zeros(1, 16, 32, 1), 1.0)
    gv = arch.global_variables()
    assert list(gv) == list(vs.vars)
    assert sum(v.numel() for v in gv.values()) == 2_561_386
    for k in vs.vars:
        assert torch.equal(gv[k].cpu(), vs.vars[k]), k


def test_net_nondefault_flags(native):
    """nr_res_blocks=2 + wider filters (BASELINE config 5), and the other nonlinearities / ungated blocks."""
    x = fo.synth_boundary(1, 32, 64, seed=2)
    for flags in (dict(nr_res_blocks=2, filter_size=16), dict(nonlinearity="elu"), dict(nonlinearity="concat_relu"),
                  dict(nonlinearity="relu", gated_res=False)):
        y = _net(native, x.astype(np.uint8), "auto", **flags)
        vs = fo.VariableStore(1234, torch.float64)
        ref = fo.inference(vs, torch.from_numpy(x).double(), 1.0, nr_res_blocks=flags.get("nr_res_blocks", 1),
                           nonlinearity=flags.get("nonlinearity", "concat_elu"),
                           gated_res=flags.get("gated_res", True), filter_size=flags.get("filter_size", 8))
        assert rel_l2(y, ref) < NET_TOL, flags


def test_net_input_dtypes_and_zero_input(native):
    """uint8 / float32 / bf16 boundaries give identical results; the reference's speed test feeds zeros."""
    x = fo.synth_boundary(2, 32, 64, seed=1)
    y8 = _net(native, x.astype(np.uint8), "auto")
    y32 = _net(native, x.astype(np.float32), "auto")
    assert torch.equal(y8, y32)
    z = _net(native, np.zeros((1, 32, 64, 1), np.float32), "auto")
    vs = fo.VariableStore(1234, torch.float64)
    ref = fo.inference(vs, torch.zeros(1, 32, 64, 1, dtype=torch.float64), 1.0)
    assert rel_l2(z, ref) < NET_TOL


def test_full_size_properties(native, golden_dir):
    """BASELINE config 2 at full size (28 cars tiled to batch 64, 128x256): determinism, batch-position
    independence, and the tcgen05 path against the CUDA-core path on the same device."""
    from utils.synthetic import load_car_boundaries
    cars = load_car_boundaries(os.path.join(golden_dir, "cars_128x256.npz"))
    idx = [i % 28 for i in range(64)]
    x = cars[idx]
    y1 = _net(native, x, "auto")
    y2 = _net(native, x, "auto")
    assert torch.equal(y1, y2)                                   # run-to-run bit-exact
    assert torch.equal(y1[0], y1[28]) and torch.equal(y1[5], y1[61])   # same car, different batch slot
    perm = torch.randperm(64, generator=torch.Generator().manual_seed(0))
    yp = _net(native, x[perm.numpy()], "auto")
    assert torch.equal(yp, y1[perm.cuda()])                      # images are independent
    ys = _net(native, x[:8], "simt")
    assert rel_l2(y1[:8], ys) < 5e-3
    assert torch.isfinite(y1).all() and y1.abs().max().item() < 1.0


def test_high_res_512x1024(native):
    """BASELINE config 4 shape (one image): fully convolutional; the tcgen05 path against the CPU oracle (fp32: its own
    error vs fp64 is 2e-7) and against the CUDA-core path."""
    x = fo.synth_boundary(1, 512, 1024, seed=0).astype(np.uint8)
    ya = _net(native, x, "auto")
    assert tuple(ya.shape) == (1, 512, 1024, 2)
    torch.set_num_threads(os.cpu_count() or 1)
    vs = fo.VariableStore(1234, torch.float32)
    with torch.no_grad():
        ref = fo.inference(vs, torch.from_numpy(x.astype(np.float32)), 1.0)
    assert rel_l2(ya, ref) < NET_TOL
    ys = _net(native, x, "simt")
    assert rel_l2(ya, ys) < 5e-3


def test_deeper_wider_variant_at_128x256_vs_oracle(native):
    """BASELINE config 5 (nr_res_blocks = 2, filter_size = 16) at the reference grid, where the 256-channel two-pass
    kernel (conv_s1w.cu) and the 128-channel skip pass are reached inside the network, against the CPU oracle."""
    x = fo.synth_boundary(2, 128, 256, seed=5).astype(np.uint8)
    y = _net(native, x, "auto", nr_res_blocks=2, filter_size=16)
    torch.set_num_threads(os.cpu_count() or 1)
    vs = fo.VariableStore(1234, torch.float32)
    with torch.no_grad():
        ref = fo.inference(vs, torch.from_numpy(x.astype(np.float32)), 1.0, nr_res_blocks=2, filter_size=16)
    assert rel_l2(y, ref) < NET_TOL


@pytest.mark.parametrize("split", [False, True], ids=["one_cta_per_image", "cluster_of_two"])
@pytest.mark.parametrize("B", [1, 5, 150])
def test_level5_fused_kernel_vs_layer_by_layer_and_oracle(native, B, split, monkeypatch):
    """csrc/conv_l5.cu: resnet_5_downsample + resnet_5_0 + up_conv_1 in one launch against the same five convolutions as
    separate launches (same operands, same bf16 rounding points, same MMA order) and against the fp64 oracle on
    bf16-rounded inputs."""
    import model.flow_architecture as arch
    monkeypatch.setattr(arch, "LEVEL5_SPLIT", split)       # csrc/conv_l5.cu or the channel-split cluster kernel csrc/conv_l5c.cu
    arch.reset_variables(77)
    g = torch.Generator().manual_seed(B)
    x = torch.randn(B, 16, 32, 64, generator=g).to(torch.bfloat16)
    xd = x.cuda()
    conc = arch.concat_elu

    def layers(xin):
        t = arch.res_block(xin, filter_size=128, nonlinearity=conc, stride=2, gated=True, name="resnet_5_downsample")
        t = arch.res_block(t, filter_size=128, nonlinearity=conc, gated=True, name="resnet_5_0")
        return arch.transpose_conv_layer(t, 3, 2, 64, "up_conv_1")

    y_ref = layers(xd)
    assert arch._level5_fusable(xd, 1, 1.0, conc, True, 128)
    y = arch._level5_fused(xd)
    torch.cuda.synchronize()
    assert native.tc_status() == 0
    assert tuple(y.shape) == (B, 16, 32, 64)
    # the cluster kernel sums every convolution's K range in two fp32 accumulators (one per issuing warp): a different summation
    # order than the per-layer kernels', i.e. a few bf16 roundings that flip and propagate through five layers
    assert rel_l2(y, y_ref) < (2e-3 if split else 1e-3)
    if B <= 5:
        vs = fo.VariableStore(77, torch.float64)
        # same creation order as the device graph above -> same variables
        xo = x.double()
        t = fo.res_block(vs, xo, filter_size=128, stride=2, gated=True, name="resnet_5_downsample")
        t = fo.res_block(vs, t, filter_size=128, gated=True, name="resnet_5_0")
        ref = fo.transpose_conv_layer(vs, t, 3, 2, 64, "up_conv_1")
        for k, v in vs.vars.items():
            assert torch.equal(arch.global_variables()[k].cpu().double(), v), k
        assert rel_l2(y, ref) < NET_TOL
    arch.reset_variables(1234)


@pytest.mark.parametrize("B", [1, 3, 80])
def test_level4_fused_kernels_vs_layer_by_layer_and_oracle(native, B, monkeypatch):
    """csrc/conv_l4.cu: resnet_4_downsample + resnet_4_0 in one launch, resnet_up_1_0 (+ nin skip) + up_conv_2 in another (a
    cluster of two CTAs per image, halo columns exchanged through distributed shared memory), against the same
    convolutions as separate launches and against the fp64 oracle on bf16-rounded inputs."""
    import model.flow_architecture as arch
    monkeypatch.setattr(arch, "LEVEL4_MAX_WAVES", 100.0)
    arch.reset_variables(78)
    g = torch.Generator().manual_seed(100 + B)
    x3 = torch.randn(B, 32, 64, 32, generator=g).to(torch.bfloat16)
    z = torch.randn(B, 16, 32, 64, generator=g).to(torch.bfloat16)
    x3d, zd = x3.cuda(), z.cuda()
    conc = arch.concat_elu

    # layer by layer, creating the variables in the order the fused wrappers do
    t = arch.res_block(x3d, filter_size=64, nonlinearity=conc, stride=2, gated=True, name="resnet_4_downsample")
    x4_ref = arch.res_block(t, filter_size=64, nonlinearity=conc, gated=True, name="resnet_4_0")
    u = arch.res_block(zd, a=x4_ref, filter_size=64, nonlinearity=conc, gated=True, name="resnet_up_1_0")
    y_ref = arch.transpose_conv_layer(u, 3, 2, 32, "up_conv_2")

    assert arch._level4_fusable(x3d, 1, 1.0, conc, True, 64)
    x4 = arch._level4_down(x3d)
    y = arch._level4_up(zd, x4_ref)
    torch.cuda.synchronize()
    assert native.tc_status() == 0
    assert tuple(x4.shape) == (B, 16, 32, 64) and tuple(y.shape) == (B, 32, 64, 32)
    assert rel_l2(x4, x4_ref) < 1e-3
    assert rel_l2(y, y_ref) < 1e-3
    if B <= 3:
        vs = fo.VariableStore(78, torch.float64)
        t = fo.res_block(vs, x3.double(), filter_size=64, stride=2, gated=True, name="resnet_4_downsample")
        x4o = fo.res_block(vs, t, filter_size=64, gated=True, name="resnet_4_0")
        uo = fo.res_block(vs, z.double(), a=x4_ref.cpu().double(), filter_size=64, gated=True, name="resnet_up_1_0")
        yo = fo.transpose_conv_layer(vs, uo, 3, 2, 32, "up_conv_2")
        for k, v in vs.vars.items():
            assert torch.equal(arch.global_variables()[k].cpu().double(), v), k
        assert rel_l2(x4, x4o) < NET_TOL
        assert rel_l2(y, yo) < NET_TOL
    arch.reset_variables(1234)


@pytest.mark.parametrize("B", [1, 5, 74])
def test_levels45_merged_launch_equals_the_three_launches(native, B):
    """csrc/conv_l4.cu MODE_ALL: levels 4 and 5 (twelve convolutions) in one launch run the very stages of the three launches
    they replace -- same operands, same order -- so the result is bit-identical, also when the launch is repeated."""
    import model.flow_architecture as arch
    arch.reset_variables(80)
    g = torch.Generator().manual_seed(300 + B)
    x3 = torch.randn(B, 32, 64, 32, generator=g).to(torch.bfloat16).cuda()
    x4_ref = arch._level4_down(x3)
    z_ref = arch._level5_fused(x4_ref)
    y_ref = arch._level4_up(z_ref, x4_ref)
    for _ in range(2):
        y, x4 = arch._level45_fused(x3)
        torch.cuda.synchronize()
        assert native.tc_status() == 0
        assert torch.equal(x4, x4_ref)
        assert torch.equal(y, y_ref)
    arch.reset_variables(1234)


def test_net_with_and_without_the_fused_level4_kernels(native):
    import model.flow_architecture as arch
    x = fo.synth_boundary(3, 128, 256, seed=13).astype(np.uint8)
    y1 = _net(native, x, "auto")
    try:
        arch.LEVEL45_MERGED = True
        y2 = _net(native, x, "auto")
        arch.LEVEL45_MERGED = False
        arch.LEVEL4_FUSED = False
        y0 = _net(native, x, "auto")
    finally:
        arch.LEVEL4_FUSED, arch.LEVEL45_MERGED = True, False
    assert torch.equal(y1, y2)               # the opt-in merged launch vs the three it replaces
    assert rel_l2(y1, y0) < 2e-3


@pytest.mark.parametrize("B", [1, 3, 80])
def test_level3_fused_kernels_vs_layer_by_layer_and_oracle(native, B, monkeypatch):
    """csrc/conv_l3.cu: conv_2 of resnet_3_downsample + resnet_3_0 in one launch, conv_2 of resnet_up_2_0 + up_conv_3 in
    another (cluster of two CTAs per image, eight tiles per CTA), against the same convolutions as separate launches and
    against the fp64 oracle on bf16-rounded inputs."""
    import model.flow_architecture as arch
    monkeypatch.setattr(arch, "LEVEL4_MAX_WAVES", 100.0)
    monkeypatch.setattr(arch, "LEVEL3_MIN_BATCH", 1)
    arch.reset_variables(79)
    g = torch.Generator().manual_seed(200 + B)
    x2 = torch.randn(B, 64, 128, 16, generator=g).to(torch.bfloat16)
    z = torch.randn(B, 32, 64, 32, generator=g).to(torch.bfloat16)
    x2d, zd = x2.cuda(), z.cuda()
    conc = arch.concat_elu

    t = arch.res_block(x2d, filter_size=32, nonlinearity=conc, stride=2, gated=True, name="resnet_3_downsample")
    x3_ref = arch.res_block(t, filter_size=32, nonlinearity=conc, gated=True, name="resnet_3_0")
    u = arch.res_block(zd, a=x3_ref, filter_size=32, nonlinearity=conc, gated=True, name="resnet_up_2_0")
    y_ref = arch.transpose_conv_layer(u, 3, 2, 16, "up_conv_3")

    assert arch._level3_fusable(x2d, 1, 1.0, conc, True, 32)
    x3 = arch._level3_down(x2d)
    y = arch._level3_up(zd, x3_ref)
    torch.cuda.synchronize()
    assert native.tc_status() == 0
    assert tuple(x3.shape) == (B, 32, 64, 32) and tuple(y.shape) == (B, 64, 128, 16)
    assert rel_l2(x3, x3_ref) < 1e-3
    assert rel_l2(y, y_ref) < 1e-3
    if B <= 3:
        vs = fo.VariableStore(79, torch.float64)
        t = fo.res_block(vs, x2.double(), filter_size=32, stride=2, gated=True, name="resnet_3_downsample")
        x3o = fo.res_block(vs, t, filter_size=32, gated=True, name="resnet_3_0")
        uo = fo.res_block(vs, z.double(), a=x3_ref.cpu().double(), filter_size=32, gated=True, name="resnet_up_2_0")
        yo = fo.transpose_conv_layer(vs, uo, 3, 2, 16, "up_conv_3")
        for k, v in vs.vars.items():
            assert torch.equal(arch.global_variables()[k].cpu().double(), v), k
        assert rel_l2(x3, x3o) < NET_TOL
        assert rel_l2(y, yo) < NET_TOL
    arch.reset_variables(1234)


def test_net_with_and_without_the_fused_level3_kernels(native, monkeypatch):
    import model.flow_architecture as arch
    x = fo.synth_boundary(3, 128, 256, seed=14).astype(np.uint8)
    y0 = _net(native, x, "auto")                                  # batch 3 is below the fused kernels' batch gate
    monkeypatch.setattr(arch, "LEVEL3_MIN_BATCH", 1)
    y1 = _net(native, x, "auto")
    assert rel_l2(y1, y0) < 2e-3


def test_net_with_and_without_the_fused_level5_kernel(native):
    import model.flow_architecture as arch
    x = fo.synth_boundary(3, 128, 256, seed=12).astype(np.uint8)
    y1 = _net(native, x, "auto")
    try:
        arch.LEVEL5_FUSED = False
        y0 = _net(native, x, "auto")
    finally:
        arch.LEVEL5_FUSED = True
    # two bf16 pipelines with different fp32 summation orders in twelve convolutions (the fused level-4 / level-5 kernels split K
    # over two accumulators), each ~3.5e-3 from the fp64 oracle (test_net_vs_golden_*): they differ by bf16 roundings that flip
    assert rel_l2(y1, y0) < 4e-3


@pytest.mark.parametrize("flags", [dict(), dict(nr_res_blocks=2, filter_size=16)], ids=["default", "deeper_wider"])
def test_net_on_planes_path_vs_raw_path_and_oracle(native, flags):
    """FLOWNET_PLANES: the same network with the activations of the 8/16/32-channel levels in MMA-ready planes form
    (csrc/conv_pl.cu, planes-writing epilogues of the head / stride-2 / transposed kernels) against the raw-tensor path
    (same kernels' arithmetic: bit-identical up to the epilogue's fused multiply-adds) and the fp64 oracle."""
    import model.flow_architecture as arch
    x = fo.synth_boundary(2, 128, 256, seed=21)
    y0 = _net(native, x.astype(np.uint8), "auto", **flags)
    try:
        arch.PLANES = True
        assert arch._planes_usable(torch.from_numpy(x.astype(np.uint8)).cuda(), 1.0, arch.concat_elu, True,
                                   flags.get("filter_size", 8))
        y1 = _net(native, x.astype(np.uint8), "auto", **flags)
    finally:
        arch.PLANES = False
    assert rel_l2(y1, y0) < 2e-3
    vs = fo.VariableStore(1234, torch.float64)
    ref = fo.inference(vs, torch.from_numpy(x).double(), 1.0, nr_res_blocks=flags.get("nr_res_blocks", 1),
                       filter_size=flags.get("filter_size", 8))
    assert rel_l2(y1, ref) < NET_TOL
    # same variables, created in the same order
    assert list(arch.global_variables()) == list(vs.vars)


def test_compiled_inference_graph_matches_eager(native):
    import model.flow_architecture as arch
    import model.flow_net as flow_net
    arch.reset_variables(1234)
    x = torch.from_numpy(fo.synth_boundary(4, 32, 64, seed=4).astype(np.uint8)).cuda()
    ye = flow_net.inference(x, 1.0).clone()
    ci = flow_net.CompiledInference(4, (32, 64))
    yg = ci(x)
    torch.cuda.synchronize()
    assert torch.equal(ye, yg)
    assert ci.launches_per_replay >= 31


# ------------------------------------------------------------------------------- small kernels
def test_nonlinearity_kernels(native):
    g = torch.Generator().manual_seed(0)
    for C in (8, 16, 3):
        x = torch.randn(5, 7, C, generator=g).to(torch.bfloat16)
        for kind, f in ((1, fo.concat_elu), (2, fo.elu), (3, fo.concat_relu), (4, torch.relu)):
            y = native.nonlinearity(x.cuda(), kind)
            ref = f(x.double())
            assert tuple(y.shape) == tuple(ref.shape)
            assert rel_l2(y, ref) < 3e-3


def test_casts(native):
    x = torch.rand(1000) * 4 - 2
    y = native.to_bf16(x.cuda())
    assert torch.equal(y.cpu(), x.to(torch.bfloat16))
    assert torch.equal(native.to_f32(y).cpu(), x.to(torch.bfloat16).float())
    u = (torch.rand(999) > 0.5).to(torch.uint8)
    assert torch.equal(native.to_bf16(u.cuda()).cpu().float(), u.float())


def test_l2_loss_kernel(native):
    g = torch.Generator().manual_seed(0)
    p = torch.randn(2, 32, 64, 2, generator=g)
    t = torch.randn(2, 32, 64, 2, generator=g)
    loss, grad = native.l2_loss(p.cuda(), t.cuda(), want_grad=True)
    ref = fo.loss_image(p.double(), t.double())
    assert abs(loss.item() - ref.item()) < 1e-5 * ref.item()
    assert torch.equal(grad.cpu(), p - t)
    assert torch.equal(native.l2_loss(p.cuda(), p.cuda())[0].cpu(), torch.zeros(1))
    e = torch.empty(0, device="cuda")
    assert native.l2_loss(e, e)[0].item() == 0.0                  # empty input


def test_adam_kernel_matches_tf1_update(native):
    g = torch.Generator().manual_seed(0)
    n = 100_003
    p0 = torch.randn(n, generator=g)
    m0 = torch.zeros(n)
    v0 = torch.zeros(n)
    pd, md, vd = p0.cuda().clone(), m0.cuda(), v0.cuda()
    po = {"a": p0.double().clone()}
    mo = {"a": m0.double().clone()}
    vo = {"a": v0.double().clone()}
    for step in (1, 2, 3):
        gr = torch.randn(n, generator=g)
        native.adam_step(pd, gr.cuda(), md, vd, step)
        fo.adam_step_tf1(po, {"a": gr.double()}, mo, vo, t=step)
    torch.cuda.synchronize()
    assert (pd.cpu().double() - po["a"]).abs().max().item() < 1e-6
    assert rel_l2(md, mo["a"]) < 1e-6 and rel_l2(vd, vo["a"]) < 1e-6


# ------------------------------------------------------------------------------- training step
def _oracle_grads(x_np, tgt_np, keep_prob=1.0, dropout_seed=None):
    """loss and autograd gradients of the fp64 oracle; keep_prob < 1 uses the CUDA path's counter-based mask
    (oracle/dropout_oracle.py) with the same per-block seeds, so both sides drop the same elements."""
    torch.set_num_threads(os.cpu_count() or 1)
    vt = fo.VariableStore(1234, torch.float64)
    xb = torch.from_numpy(x_np).double()
    fo.inference(vt, xb[:1, :16, :32], 1.0)
    vt.vars = {k: v.clone().requires_grad_(True) for k, v in vt.vars.items()}
    loss = fo.loss_image(fo.inference(vt, xb, keep_prob, dropout_seed=dropout_seed), torch.from_numpy(tgt_np).double())
    loss.backward()
    return loss.item(), vt


GRAD_TOL = 2e-2      # BASELINE.md section 4 / SURVEY 8d: per-variable gradient rel-L2, bf16 activations and activation gradients


def _check_adam_against_tf1(flat, before, lr, t):
    """TF1 ApplyAdam restated in fp64 (oracle adam_step_tf1: eps OUTSIDE the square root, lr_t folded) on the gradient
    the device itself produced: every entry of the parameter update and of both slots, not just the clear ones."""
    p0, m0, v0 = before
    g = flat.grad.double().cpu()
    pd, md, vd = {"a": p0.clone()}, {"a": m0.clone()}, {"a": v0.clone()}
    fo.adam_step_tf1(pd, {"a": g}, md, vd, t=t, lr=lr)
    upd_ref = pd["a"] - p0
    upd_got = flat.param.double().cpu() - p0
    assert rel_l2(upd_got, upd_ref) < 1e-3                      # fp32 kernel vs fp64 restatement
    assert (upd_got - upd_ref).abs().max().item() < 1e-3 * lr + 1e-9
    assert rel_l2(flat.m, md["a"]) < 1e-6 and rel_l2(flat.v, vd["a"]) < 1e-6


def _training_step_case(native, x, tgt, keep_prob, seed):
    """one or two fwd+bwd+Adam steps on the device vs oracle autograd: loss, every gradient, the Adam update"""
    import model.flow_architecture as arch
    import model.flow_net as flow_net
    arch.reset_variables(1234)
    flow_net.reset_training_state()
    try:
        flow_net.begin_training()
        worst = 0.0
        for t in (1, 2):
            p = flow_net.inference(torch.from_numpy(x).cuda(), keep_prob, dropout_seed=seed)
            err = flow_net.loss_image_train(p, torch.from_numpy(tgt).cuda())
            flat = flow_net.ensure_training_state()
            before = (flat.param.double().cpu(), flat.m.double().cpu(), flat.v.double().cpu())
            loss = flow_net.train(err, 1e-4)
            torch.cuda.synchronize()
            assert native.tc_status() == 0
            _check_adam_against_tf1(flat, before, 1e-4, t)
            if t == 1:
                ref_loss, vt = _oracle_grads(x.astype(np.float64), tgt.astype(np.float64), keep_prob, seed)
                assert abs(loss.item() - ref_loss) < 2e-2 * abs(ref_loss)
                assert abs(err.item() - ref_loss) < 2e-2 * abs(ref_loss)
                for name, v in vt.vars.items():
                    e = rel_l2(flat.grad_view(name), v.grad)
                    worst = max(worst, e)
                    assert e < GRAD_TOL, (name, e)
        print("worst per-variable gradient rel-L2:", worst)
        return worst
    finally:
        flow_net.end_training()
        flow_net.reset_training_state()
        arch.reset_variables(1234)


def test_training_step_with_dropout_vs_oracle_same_mask(native):
    """keep_prob = 0.7 -- the reference's training setting (train/flow_train.py:21; flow_architecture.py:144-145).  The
    device mask is a counter-based hash; oracle/dropout_oracle.py restates it, so the oracle's autograd runs the SAME
    masked network: forward loss, every gradient (mask in the forward transform, the recomputation, the weight-gradient
    transform and the prologue adjoint) and the Adam update."""
    x = fo.synth_boundary(2, 32, 64, seed=3).astype(np.uint8)
    g = torch.Generator().manual_seed(5)
    tgt = (torch.rand(2, 32, 64, 2, generator=g) * 1.6 - 0.8).numpy().astype(np.float32)
    _training_step_case(native, x, tgt, 0.7, seed=11)


def test_dropout_forward_vs_oracle_same_mask(native):
    import model.flow_architecture as arch
    import model.flow_net as flow_net
    x = fo.synth_boundary(2, 48, 80, seed=8)
    arch.reset_variables(1234)
    y = flow_net.inference(torch.from_numpy(x.astype(np.uint8)).cuda(), 0.7, dropout_seed=21)
    vs = fo.VariableStore(1234, torch.float64)
    ref = fo.inference(vs, torch.from_numpy(x).double(), 0.7, dropout_seed=21)
    ref_other = fo.inference(vs, torch.from_numpy(x).double(), 0.7, dropout_seed=22)
    assert rel_l2(y, ref) < NET_TOL
    assert rel_l2(y, ref_other) > 10 * NET_TOL            # a different mask is a different network


def test_training_step_128x256_vs_oracle(native):
    """the reference grid: the persistent 128x256-level wgrad / dgrad paths inside a whole step (2 images)"""
    x = fo.synth_boundary(2, 128, 256, seed=9).astype(np.uint8)
    g = torch.Generator().manual_seed(6)
    tgt = (torch.rand(2, 128, 256, 2, generator=g) * 1.6 - 0.8).numpy().astype(np.float32)
    _training_step_case(native, x, tgt, 1.0, seed=0)


def test_backward_primitives(native):
    g = torch.Generator().manual_seed(0)
    # gate
    c2 = torch.randn(3, 5, 7, 16, generator=g, dtype=torch.float64).to(torch.bfloat16)
    gg = torch.randn(3, 5, 7, 8, generator=g, dtype=torch.float64).to(torch.bfloat16)
    c2r = c2.double().requires_grad_(True)
    (c2r[..., :8] * torch.sigmoid(c2r[..., 8:]) * gg.double()).sum().backward()
    assert rel_l2(native.gate_bwd(c2.cuda(), gg.cuda()), c2r.grad) < 5e-3
    # prologues
    for kind, f in ((1, fo.concat_elu), (2, fo.elu), (3, fo.concat_relu), (4, torch.relu)):
        x = torch.randn(2, 6, 5, 8, generator=g, dtype=torch.float64).to(torch.bfloat16)
        mult = 2 if kind in (1, 3) else 1
        dA = torch.randn(2, 6, 5, 8 * mult, generator=g, dtype=torch.float64).to(torch.bfloat16)
        xr = x.double().requires_grad_(True)
        (f(xr) * dA.double()).sum().backward()
        assert rel_l2(native.prologue_bwd(dA.cuda(), x.cuda(), kind), xr.grad) < 5e-3
    # shortcut (pooled + front pad, odd size)
    x = torch.randn(2, 7, 9, 8, generator=g, dtype=torch.float64)
    gr = torch.randn(2, 4, 5, 16, generator=g, dtype=torch.float64).to(torch.bfloat16)
    xr = x.clone().requires_grad_(True)
    sc = torch.nn.functional.pad(fo.avg_pool_2x2_same(xr), (8, 0))
    (sc * gr.double()).sum().backward()
    assert rel_l2(native.shortcut_bwd(gr.cuda(), (2, 7, 9, 8), True), xr.grad) < 5e-3
    # bias grad
    dY = torch.randn(4, 9, 11, 32, generator=g, dtype=torch.float64).to(torch.bfloat16)
    out = torch.empty(32, device="cuda")
    native.bias_grad(dY.cuda(), out)
    assert rel_l2(out, dY.double().sum(dim=(0, 1, 2))) < 1e-5


@pytest.mark.parametrize("ksize,stride,pro,H,W,C,N", [(3, 1, 1, 12, 20, 8, 16), (3, 2, 1, 16, 24, 16, 32), (1, 1, 1, 8, 8, 16, 8),
                                                      (3, 1, 0, 9, 13, 1, 8), (3, 2, 2, 15, 17, 8, 8)])
def test_conv_wgrad_vs_autograd(native, ksize, stride, pro, H, W, C, N):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, H, W, C, generator=g, dtype=torch.float64).to(torch.bfloat16)
    mult = 2 if pro in (1, 3) else 1
    w = torch.randn(ksize, ksize, C * mult, N, generator=g, dtype=torch.float64).requires_grad_(True)
    OH, OW = -(-H // stride), -(-W // stride)
    dY = torch.randn(3, OH, OW, N, generator=g, dtype=torch.float64).to(torch.bfloat16)
    a = bf16_round(PRO_FN[pro](x.double())) if pro else x.double()
    (fo.conv2d_same(a, w, stride) * dY.double()).sum().backward()
    out = torch.empty(ksize * ksize, C * mult, N, device="cuda")
    native.conv_wgrad(x.cuda(), dY.cuda(), out, ksize=ksize, stride=stride, prologue=pro)
    assert rel_l2(out, w.grad.reshape(ksize * ksize, C * mult, N)) < 1e-4


@pytest.mark.parametrize("impl", ["simt", "tc"])
@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 5, 6, 16, 8), (3, 16, 32, 32, 16), (2, 8, 16, 128, 64), (5, 32, 64, 16, 8)])
def test_tconv_wgrad_vs_autograd(native, impl, B, H, W, Cin, Cout):
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, H, W, Cin, generator=g, dtype=torch.float64).to(torch.bfloat16)
    w = torch.randn(3, 3, Cout, Cin, generator=g, dtype=torch.float64).requires_grad_(True)     # TF [k,k,Cout,Cin]
    dY = torch.randn(B, 2 * H, 2 * W, Cout, generator=g, dtype=torch.float64).to(torch.bfloat16)
    (fo.conv2d_transpose_same(x.double(), w, 2) * dY.double()).sum().backward()
    out = torch.full((9, Cout, Cin), float("nan"), device="cuda")
    native.tconv_wgrad(x.cuda(), dY.cuda(), out, impl=native.IMPL_TCGEN05 if impl == "tc" else native.IMPL_SIMT)
    torch.cuda.synchronize()
    assert native.tc_status() == 0
    assert rel_l2(out.view(3, 3, Cout, Cin), w.grad) < 1e-4


WGRAD_TC_CASES = [
    # (B, H, W, C, N, ksize, stride, pro)
    (2, 16, 32, 8, 16, 3, 1, 1), (3, 24, 40, 8, 8, 3, 1, 1), (2, 16, 32, 16, 32, 3, 1, 1), (2, 16, 32, 32, 64, 3, 1, 1),
    (2, 16, 32, 64, 128, 3, 1, 1), (3, 8, 16, 128, 256, 3, 1, 1), (3, 8, 16, 128, 128, 3, 1, 1), (2, 13, 21, 16, 16, 3, 1, 1),
    (2, 32, 64, 8, 16, 3, 2, 1), (2, 16, 32, 64, 128, 3, 2, 1), (1, 18, 36, 16, 32, 3, 2, 1),
    (2, 16, 32, 64, 64, 1, 1, 1), (2, 16, 32, 8, 8, 1, 1, 1),
    (2, 16, 32, 8, 8, 3, 1, 0), (2, 16, 32, 32, 16, 3, 1, 2), (2, 16, 32, 16, 16, 3, 1, 3), (2, 16, 32, 16, 16, 3, 1, 4),
    (40, 32, 64, 8, 16, 3, 1, 1),
]


@pytest.mark.parametrize("B,H,W,C,N,ksize,stride,pro", WGRAD_TC_CASES)
def test_conv_wgrad_tcgen05_vs_autograd(native, B, H, W, C, N, ksize, stride, pro):
    """wgrad_tc.cu (pixels are the contraction dimension, MN-major operands) against torch autograd in fp64
    on the bf16-rounded operands, and against the CUDA-core kernel."""
    g = torch.Generator().manual_seed(11)
    x = bf16_round(torch.randn(B, H, W, C, generator=g, dtype=torch.float64))
    OH, OW = -(-H // stride), -(-W // stride)
    dY = bf16_round(torch.randn(B, OH, OW, N, generator=g, dtype=torch.float64))
    mult = 2 if pro in (1, 3) else 1
    w = torch.zeros(ksize, ksize, C * mult, N, dtype=torch.float64, requires_grad=True)
    a = bf16_round(PRO_FN[pro](x)) if pro else x
    (fo.conv2d_same(a, w, stride) * dY).sum().backward()
    outs = []
    for impl in (native.IMPL_TCGEN05, native.IMPL_SIMT):
        out = torch.full((ksize * ksize, C * mult, N), float("nan"), device="cuda")
        native.conv_wgrad(x.to("cuda", torch.bfloat16), dY.to("cuda", torch.bfloat16), out, ksize=ksize, stride=stride,
                          prologue=pro, impl=impl)
        torch.cuda.synchronize()
        assert native.tc_status() == 0
        outs.append(out)
    ref = w.grad.reshape(ksize * ksize, C * mult, N)
    assert rel_l2(outs[0], ref) < 1e-4
    assert rel_l2(outs[0], outs[1]) < 1e-4


def test_first_conv_wgrad_kernel_vs_cuda_core(native):
    """Cin = 1 (the boundary conv): dedicated reduction kernel vs the generic CUDA-core one."""
    g = torch.Generator().manual_seed(13)
    x = (torch.rand(5, 32, 64, 1, generator=g) > 0.7).to("cuda", torch.bfloat16)
    dY = torch.randn(5, 32, 64, 8, generator=g).to("cuda", torch.bfloat16)
    outs = []
    for impl in (native.IMPL_AUTO, native.IMPL_SIMT):
        out = torch.full((9, 1, 8), float("nan"), device="cuda")
        native.conv_wgrad(x, dY, out, ksize=3, stride=1, prologue=0, impl=impl)
        outs.append(out)
    assert rel_l2(outs[0], outs[1]) < 1e-5


def test_conv_wgrad_tcgen05_dropout_mask_matches_cuda_core_kernel(native):
    g = torch.Generator().manual_seed(12)
    x = torch.randn(2, 16, 32, 16, generator=g).to("cuda", torch.bfloat16)
    dY = torch.randn(2, 16, 32, 32, generator=g).to("cuda", torch.bfloat16)
    outs = []
    for impl in (native.IMPL_TCGEN05, native.IMPL_SIMT):
        out = torch.empty(9, 32, 32, device="cuda")
        native.conv_wgrad(x, dY, out, ksize=3, stride=1, prologue=1, keep_prob=0.7, dropout_seed=5, impl=impl)
        outs.append(out)
    # same keep mask; the tensor-core operand is rounded to bf16 AFTER the 1/keep_prob scaling (2^-9 per element),
    # a wrong mask would differ by O(1)
    assert rel_l2(outs[0], outs[1]) < 5e-3


def test_training_step_vs_oracle(native, golden_dir):
    """One fwd+bwd+Adam step on the golden 2x32x64 case: loss, every gradient and the updated variables."""
    import model.flow_architecture as arch
    import model.flow_net as flow_net
    z = np.load(os.path.join(golden_dir, "golden_train.npz"))
    x = fo.synth_boundary(2, 32, 64, seed=3).astype(np.uint8)
    tgt = z["target"].astype(np.float32)
    arch.reset_variables(1234)
    flow_net.reset_training_state()
    try:
        flow_net.begin_training()
        p = flow_net.inference(torch.from_numpy(x).cuda(), 1.0)
        err = flow_net.loss_image_train(p, torch.from_numpy(tgt).cuda())
        flat = flow_net.ensure_training_state()
        before = (flat.param.double().cpu(), flat.m.double().cpu(), flat.v.double().cpu())
        loss = flow_net.train(err, 1e-4)
        torch.cuda.synchronize()
        assert native.tc_status() == 0
        ref_loss, vt = _oracle_grads(x.astype(np.float64), z["target"])
        assert abs(ref_loss - float(z["loss"])) < 1e-9 * abs(ref_loss)
        assert abs(loss.item() - ref_loss) < 2e-2 * abs(ref_loss)
        assert abs(err.item() - ref_loss) < 2e-2 * abs(ref_loss)
        worst = 0.0
        for name, v in vt.vars.items():
            e = rel_l2(flat.grad_view(name), v.grad)
            worst = max(worst, e)
            assert e < GRAD_TOL, (name, e)       # bf16 activations and activation gradients
        print("worst per-variable gradient rel-L2:", worst)
        # TF1 Adam on the device's own gradient: every entry of the update and of both slots
        _check_adam_against_tf1(flat, before, 1e-4, 1)
    finally:
        flow_net.end_training()
        flow_net.reset_training_state()
        arch.reset_variables(1234)


def test_training_loss_decreases(native):
    import model.flow_architecture as arch
    import model.flow_net as flow_net
    arch.reset_variables(1234)
    flow_net.reset_training_state()
    b, s = flow_net.inputs(4, shape=(32, 64), seed=0)
    try:
        flow_net.begin_training()
        losses = []
        for _ in range(8):
            p = flow_net.inference(b, 1.0)
            err = flow_net.loss_image_train(p, s)
            losses.append(flow_net.train(err, 1e-3).item())
        assert all(np.isfinite(losses))
        assert losses[-1] < 0.7 * losses[0], losses
    finally:
        flow_net.end_training()
        flow_net.reset_training_state()
        arch.reset_variables(1234)


def test_stream_host_matches_eager(native):
    """The overlapped serving loop (pinned host in/out, side-stream copies) returns the same fields."""
    import model.flow_architecture as arch
    import model.flow_net as flow_net
    arch.reset_variables(1234)
    ci = flow_net.CompiledInference(2, (32, 64))
    xs = [torch.from_numpy(fo.synth_boundary(2, 32, 64, seed=10 + i).astype(np.uint8)).pin_memory() for i in range(5)]
    outs = [torch.empty(2, 32, 64, 2).pin_memory() for _ in range(5)]
    ci.stream_host(xs, outs)
    torch.cuda.synchronize()
    for x, o in zip(xs, outs):
        ref = flow_net.inference(x.cuda(), 1.0)
        assert torch.equal(ref.cpu(), o)


def test_checkpoint_resume_continues_bit_exactly(native, tmp_path):
    """utils/checkpoint.py: stop after 2 steps, restore variables + Adam slots into a fresh
    graph, continue -- same variables as the uninterrupted run (every reduction in the step is ordered)."""
    import model.flow_architecture as arch
    import model.flow_net as flow_net
    from utils import checkpoint

    def steps(k0, k1):
        for k in range(k0, k1):
            b, s = flow_net.inputs(2, shape=(32, 64), seed=10 * k)
            p = flow_net.inference(b, 1.0)
            err = flow_net.loss_image_train(p, s)
            flow_net.train(err, 1e-3)

    try:
        arch.reset_variables(1234)
        flow_net.reset_training_state()
        flow_net.begin_training()
        steps(0, 2)
        path = checkpoint.save(str(tmp_path), 2, arch.global_variables(), flow_net.training_state())
        steps(2, 4)
        want = {k: v.clone() for k, v in arch.global_variables().items()}
        flow_net.end_training()

        arch.reset_variables(77)                     # different random init: everything must come from the file
        flow_net.reset_training_state()
        assert checkpoint.restore(path, arch.VARIABLES) == 2
        flow_net.begin_training()
        b, _ = flow_net.inputs(2, shape=(32, 64), seed=0)
        flow_net.inference(b, 1.0)                   # creates the variables (from the preloaded values)
        arch.TAPE.clear()
        checkpoint.restore(path, arch.VARIABLES, flow_net.ensure_training_state(), strict=False)
        steps(2, 4)
        got = arch.global_variables()
        assert set(got) == set(want)
        for k in want:
            assert torch.equal(got[k], want[k]), k
    finally:
        flow_net.end_training()
        flow_net.reset_training_state()
        arch.reset_variables(1234)


def test_boundary_gradient_vs_oracle_autograd(native, golden_dir):
    """d loss / d boundary (the first conv's data gradient + the front-padded shortcut)."""
    import model.flow_architecture as arch
    import model.flow_net as flow_net
    z = np.load(os.path.join(golden_dir, "golden_train.npz"))
    x = fo.synth_boundary(2, 32, 64, seed=3).astype(np.float64)
    tgt = z["target"]
    vs = fo.VariableStore(1234, torch.float64)
    xt = torch.from_numpy(x).requires_grad_(True)
    ref_loss = fo.loss_image(fo.inference(vs, xt, 1.0), torch.from_numpy(tgt))
    ref_loss.backward()
    arch.reset_variables(1234)
    loss, g = flow_net.boundary_gradient(torch.from_numpy(x.astype(np.float32)).cuda(),
                                         torch.from_numpy(tgt.astype(np.float32)).cuda())
    torch.cuda.synchronize()
    assert native.tc_status() == 0
    assert g.shape == (2, 32, 64, 1) and g.dtype == torch.float32
    assert abs(loss.item() - ref_loss.item()) < 2e-2 * abs(ref_loss.item())
    assert rel_l2(g, xt.grad) < 5e-2


def test_training_from_tfrecords_queue(native, tmp_path):
    """dataset file -> native parser -> shuffle queue -> pinned uint8/float32 -> GPU -> train step."""
    import model.flow_architecture as arch
    import model.flow_net as flow_net
    import input.flow_input as fi
    from utils.synthetic import synth_boundary, synth_sflow
    b = synth_boundary(12, 32, 64, seed=5)
    s = synth_sflow(b)
    path = str(tmp_path / "train.tfrecords")
    fi.write_tfrecords(path, b, s)
    pairs = {s[i].tobytes(): b[i].tobytes() for i in range(12)}
    try:
        arch.reset_variables(1234)
        flow_net.reset_training_state()
        flow_net.begin_training()
        losses = []
        for _ in range(3):
            bd, sd = flow_net.inputs(4, shape=(32, 64), data_glob=path)
            assert bd.is_cuda and bd.dtype == torch.uint8 and sd.dtype == torch.float32
            for j in range(4):
                assert pairs[sd[j].cpu().numpy().tobytes()] == bd[j].cpu().numpy().tobytes()
            p = flow_net.inference(bd, 1.0)
            err = flow_net.loss_image_train(p, sd)
            flow_net.train(err, 1e-3)
            losses.append(err.item())
        assert all(np.isfinite(losses))
    finally:
        flow_net.end_training()
        flow_net.reset_training_state()
        arch.reset_variables(1234)
        for q in flow_net._QUEUES.values():
            q.close()
        flow_net._QUEUES.clear()


@pytest.mark.parametrize("B,H,W,C,gate", [(2, 32, 64, 8, True), (2, 32, 64, 16, True), (2, 32, 32, 32, True), (2, 16, 32, 64, True),
                                          (3, 8, 16, 128, True), (2, 32, 64, 16, False)])
def test_dropout_prologue_tcgen05_matches_cuda_core_mask(native, B, H, W, C, gate):
    """tf.nn.dropout on concat_elu(x_1) (flow_architecture.py:144-145) inside the persistent tcgen05 kernels: the same
    keep mask as the CUDA-core kernel (a different mask would differ by O(1)); the tensor-core operand is rounded to
    bf16 after the 1/keep_prob scaling, hence the looser bound."""
    g = torch.Generator().manual_seed(21)
    N = 2 * C if gate else C
    x = torch.randn(B, H, W, C, generator=g).to("cuda", torch.bfloat16)
    w16 = ((torch.rand(9, 2 * C, N, generator=g) * 2 - 1) * (6.0 / (9 * (2 * C + N))) ** 0.5).to("cuda", torch.bfloat16)
    bias = torch.rand(N, generator=g).to("cuda")
    res = torch.randn(B, H, W, C, generator=g).to("cuda", torch.bfloat16) if gate else None
    w_tc = native.pack_conv_weights_tc(w16, None, 3, 2 * C, 0, N, gate)
    ys = []
    for impl in (native.IMPL_TCGEN05, native.IMPL_SIMT):
        y = torch.empty((B, H, W, C), device="cuda", dtype=torch.bfloat16)
        native.conv_fwd(x, w16, bias, y, B=B, H=H, W=W, Cin=C, Cout=N, ksize=3, stride=1, prologue=1,
                        epilogue=1 if gate else 0, impl=impl, w_tc=w_tc, res=res, keep_prob=0.7, dropout_seed=9)
        torch.cuda.synchronize()
        assert native.tc_status() == 0
        assert native.last_used_tcgen05() == (impl == native.IMPL_TCGEN05)
        ys.append(y)
    assert rel_l2(ys[0], ys[1]) < 1e-2
    y1 = torch.empty_like(ys[0])
    native.conv_fwd(x, w16, bias, y1, B=B, H=H, W=W, Cin=C, Cout=N, ksize=3, stride=1, prologue=1, epilogue=1 if gate else 0,
                    impl=native.IMPL_TCGEN05, w_tc=w_tc, res=res, keep_prob=0.7, dropout_seed=10)
    assert rel_l2(y1, ys[0]) > 0.05              # another seed, another mask


@pytest.mark.parametrize("B,H,W,C,keep", [(2, 32, 64, 8, 1.0), (2, 32, 64, 16, 0.7), (2, 32, 32, 32, 1.0), (2, 16, 32, 64, 0.7),
                                          (3, 8, 16, 128, 1.0), (2, 20, 28, 8, 1.0)])
def test_gate_backward_epilogue_vs_two_step(native, B, H, W, C, keep):
    """FN_EPI_GATE_BWD (recompute [u|v], gate adjoint in the epilogue) against conv(EPI_BIAS) + fn_gate_bwd."""
    g = torch.Generator().manual_seed(31)
    N = 2 * C
    x = torch.randn(B, H, W, C, generator=g).to("cuda", torch.bfloat16)
    w16 = ((torch.rand(9, 2 * C, N, generator=g) * 2 - 1) * (6.0 / (9 * (2 * C + N))) ** 0.5).to("cuda", torch.bfloat16)
    bias = (torch.rand(N, generator=g) - 0.5).to("cuda")
    gy = torch.randn(B, H, W, C, generator=g).to("cuda", torch.bfloat16)
    w_tc = native.pack_conv_weights_tc(w16, None, 3, 2 * C, 0, N, True)
    kw = dict(B=B, H=H, W=W, Cin=C, Cout=N, ksize=3, stride=1, prologue=1, w_tc=w_tc, keep_prob=keep, dropout_seed=4)
    assert native.conv_fwd(x, w16, bias, x, epilogue=native.EPI_GATE_BWD, res=gy, query=True, **kw)
    fused = torch.full((B, H, W, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    native.conv_fwd(x, w16, bias, fused, epilogue=native.EPI_GATE_BWD, res=gy, impl=native.IMPL_TCGEN05, **kw)
    c2 = torch.empty((B, H, W, N), device="cuda", dtype=torch.bfloat16)
    native.conv_fwd(x, w16, bias, c2, epilogue=native.EPI_BIAS, impl=native.IMPL_SIMT, **kw)
    two = native.gate_bwd(c2, gy)
    torch.cuda.synchronize()
    assert native.tc_status() == 0
    # the two-step path rounds [u|v] to bf16 before the gate adjoint, the fused one does not
    assert rel_l2(fused, two) < (1.2e-2 if keep < 1.0 else 8e-3)
    with pytest.raises(ValueError):
        native.conv_fwd(x, w16, bias, fused, epilogue=native.EPI_GATE_BWD, res=gy, impl=native.IMPL_SIMT, **kw)


def test_dropout_mask_statistics_and_scalar_vector_agreement(native):
    """keep fraction ~ keep_prob, kept values scaled by 1/keep_prob (tf.nn.dropout, flow_architecture.py:144-145); the
    8-wide and the scalar mask code (C = 8 vs C = 4 views of the same element stream) agree element by element."""
    n = 1 << 16
    ones8 = torch.ones(n, 8, device="cuda", dtype=torch.bfloat16)
    m8 = native.prologue_bwd(ones8, ones8, 0, keep_prob=0.7, seed=3).float()
    vals = set(m8.unique().tolist())
    assert vals == {0.0, float(torch.tensor(1 / 0.7).to(torch.bfloat16))}
    frac = (m8 > 0).float().mean().item()
    assert abs(frac - 0.7) < 5e-3
    ones4 = torch.ones(2 * n, 4, device="cuda", dtype=torch.bfloat16)          # same flat element indices, scalar kernel
    m4 = native.prologue_bwd(ones4, ones4, 0, keep_prob=0.7, seed=3).float()
    assert torch.equal(m4.reshape(-1), m8.reshape(-1))
    m8b = native.prologue_bwd(ones8, ones8, 0, keep_prob=0.7, seed=4).float()
    assert 0.35 < (m8b.reshape(-1) == m8.reshape(-1)).float().mean().item() < 0.75   # independent masks agree on ~58 %


@pytest.mark.parametrize("B,H,W,C,N,kind,acc,keep", [(2, 32, 64, 16, 16, 1, False, 1.0), (2, 32, 64, 8, 16, 1, True, 1.0),
                                                     (2, 32, 32, 32, 64, 1, True, 0.7), (2, 20, 28, 64, 64, 1, False, 0.7),
                                                     (2, 32, 64, 16, 32, 3, True, 1.0), (40, 32, 64, 16, 16, 1, True, 1.0)])
def test_prologue_adjoint_epilogue_vs_two_step(native, B, H, W, C, N, kind, acc, keep):
    """FN_EPI_PRO_BWD (data-gradient conv + prologue adjoint + dropout mask in one launch) against the conv followed
    by fn_prologue_bwd."""
    g = torch.Generator().manual_seed(41)
    mult = 2 if kind in (1, 3) else 1
    cy = N // mult
    gx = torch.randn(B, H, W, C, generator=g).to("cuda", torch.bfloat16)
    w16 = ((torch.rand(9, C, N, generator=g) * 2 - 1) * 0.2).to("cuda", torch.bfloat16)
    z = torch.randn(B, H, W, cy, generator=g).to("cuda", torch.bfloat16)
    base = torch.randn(B, H, W, cy, generator=g).to("cuda", torch.bfloat16)
    bias = torch.zeros(N, device="cuda")
    w_tc = native.pack_conv_weights_tc(w16, None, 3, C, 0, N, False)
    kw = dict(B=B, H=H, W=W, Cin=C, Cout=N, ksize=3, stride=1, prologue=0, w_tc=w_tc)
    fused = base.clone()
    akw = dict(epilogue=native.EPI_PRO_BWD, res=z, keep_prob=keep, dropout_seed=6, adj_prologue=kind, adj_accumulate=acc)
    assert native.conv_fwd(gx, w16, bias, fused, query=True, **kw, **akw)
    native.conv_fwd(gx, w16, bias, fused, impl=native.IMPL_TCGEN05, **kw, **akw)
    dA = torch.empty((B, H, W, N), device="cuda", dtype=torch.bfloat16)
    native.conv_fwd(gx, w16, bias, dA, epilogue=native.EPI_BIAS, impl=native.IMPL_TCGEN05, **kw)
    two = native.prologue_bwd(dA, z, kind, out=base.clone(), accumulate=acc, keep_prob=keep, seed=6)
    torch.cuda.synchronize()
    assert native.tc_status() == 0
    assert rel_l2(fused, two) < 2e-3


@pytest.mark.parametrize("B,H,W,C,kind,acc", [(3, 64, 128, 16, 1, True), (2, 32, 64, 32, 1, False), (2, 16, 32, 64, 1, True),
                                              (2, 19, 27, 16, 3, True), (70, 16, 16, 32, 1, True), (150, 16, 16, 64, 3, False)])
def test_transposed_conv_prologue_adjoint_vs_two_step(native, B, H, W, C, kind, acc):
    """fn_tconv_fwd with FN_EPI_PRO_BWD (data gradient of a stride-2 conv + the prologue's adjoint in one launch) against
    the transposed conv followed by fn_prologue_bwd."""
    g = torch.Generator().manual_seed(43)
    gx = torch.randn(B, H, W, C, generator=g).to("cuda", torch.bfloat16)
    w16 = ((torch.rand(9, C, C, generator=g) * 2 - 1) * 0.2).to("cuda", torch.bfloat16)
    z = torch.randn(B, 2 * H, 2 * W, C // 2, generator=g).to("cuda", torch.bfloat16)
    base = torch.randn(B, 2 * H, 2 * W, C // 2, generator=g).to("cuda", torch.bfloat16)
    bias = torch.zeros(C, device="cuda")
    w_tc = native.pack_conv_weights_tc(w16, None, 3, C, 0, C, False)
    kw = dict(B=B, H=H, W=W, Cin=C, Cout=C, w_tc=w_tc, impl=native.IMPL_TCGEN05)
    akw = dict(epilogue=native.EPI_PRO_BWD, res=z, adj_prologue=kind, adj_accumulate=acc)
    fused = base.clone()
    assert native.tconv_fwd(gx, w16, bias, fused, query=True, **kw, **akw)
    native.tconv_fwd(gx, w16, bias, fused, **kw, **akw)
    dA = torch.empty((B, 2 * H, 2 * W, C), device="cuda", dtype=torch.bfloat16)
    native.tconv_fwd(gx, w16, bias, dA, **kw)
    two = native.prologue_bwd(dA, z, kind, out=base.clone(), accumulate=acc)
    torch.cuda.synchronize()
    assert native.tc_status() == 0
    assert rel_l2(fused, two) < 2e-3
    # no kernel for it: the query says so and the call refuses instead of dropping the adjoint
    assert not native.tconv_fwd(gx, w16, bias, fused, query=True, **kw, epilogue=native.EPI_PRO_BWD, res=z, adj_prologue=native.PRO_ELU)
    with pytest.raises(ValueError):
        native.tconv_fwd(gx, w16, bias, fused, **kw, epilogue=native.EPI_PRO_BWD, res=z, adj_prologue=native.PRO_ELU)


def test_flow_test_script_on_a_trained_checkpoint(native, tmp_path, monkeypatch):
    """test/flow_test.py (reference test/flow_test.py:52-100): restores the flag-named checkpoint, runs image by image,
    builds the [true | generated | difference] display image.  Uses a TFRecords test set written on the spot."""
    import importlib.util
    import model.flow_architecture as arch
    import model.flow_net as flow_net
    import input.flow_input as fi
    from model import flags
    from utils import checkpoint
    from utils.experiment_manager import make_checkpoint_path
    from utils.synthetic import synth_boundary, synth_sflow
    spec = importlib.util.spec_from_file_location(
        "flow_test_script", os.path.join(os.path.dirname(os.path.abspath(native.__file__)), "test", "flow_test.py"))
    ft = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ft)
    b = synth_boundary(3, 128, 256, seed=8)
    s = synth_sflow(b)
    rec = str(tmp_path / "test.tfrecords")
    fi.write_tfrecords(rec, b, s)
    old = {k: getattr(flags.FLAGS, k) for k in ("base_dir", "data_glob", "out_dir")}
    try:
        flags.FLAGS.base_dir = str(tmp_path / "ckpt")
        flags.FLAGS.data_glob = rec
        flags.FLAGS.out_dir = str(tmp_path / "out")
        monkeypatch.chdir(tmp_path)                  # no ../data/computed_car_flow here
        arch.reset_variables(4321)
        flow_net.inference(torch.from_numpy(b[:1]).cuda(), 1.0)          # create the variables of a "trained" net
        want = flow_net.inference(torch.from_numpy(b).cuda(), 1.0).cpu().numpy()
        checkpoint.save(make_checkpoint_path(flags.FLAGS.base_dir, flags.FLAGS), 7, arch.global_variables())
        arch.reset_variables(1234)
        results = ft.evaluate()
        assert len(results) == 3
        for i, (name, err) in enumerate(results):
            ref = np.linalg.norm(s[i] - want[i]) / np.linalg.norm(s[i])
            assert abs(err - ref) < 1e-3 * ref, (name, err, ref)        # the restored net produced these fields
            img = np.load(os.path.join(flags.FLAGS.out_dir, "flow_%04d.npy" % i))
            assert img.shape == (128, 3 * 256)
    finally:
        for k, v in old.items():
            setattr(flags.FLAGS, k, v)
        arch.reset_variables(1234)


# ---------------------------------------------------------------- lattice-Boltzmann data generator
def _lbm_case(B, ny, nx, seed):
    rng = np.random.default_rng(seed)
    solid = np.zeros((B, ny, nx), np.uint8)
    solid[:, 0, :] = solid[:, -1, :] = 1
    for b in range(B):
        cy, cx = rng.integers(6, ny - 6), rng.integers(8, nx // 2)
        yy, xx = np.mgrid[0:ny, 0:nx]
        solid[b][(yy - cy) ** 2 + (xx - cx) ** 2 < rng.integers(3, 6) ** 2] = 1
    return solid


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-11), (torch.float32, 2e-4)])
def test_lbm_kernel_vs_oracle(native, dtype, tol):
    """csrc/lbm.cu against oracle/lbm_oracle.py: same operations in the same order, so fp64 agrees to rounding."""
    from oracle import lbm_oracle as lo
    from fluid_generator import lbm
    solid = _lbm_case(2, 24, 40, seed=1)
    steps = 60
    sim = lbm.ChannelFlow(solid, dtype=dtype)
    sim.step(steps)
    vel, rho = sim.fields(want_rho=True)
    torch.cuda.synchronize()
    r, vx, vy = lo.run(solid.astype(bool), steps)
    ref = np.stack([vx, vy], axis=-1)
    assert np.isfinite(vel.cpu().numpy()).all()
    if dtype == torch.float64:      # fields() rounds to fp32 on output: compare the populations' moments in fp64
        f = sim.f_a.cpu().numpy().transpose(1, 0, 2, 3).copy()
        lo.apply_bc(f, lo.inlet_profile(24, 0.1), 1.0)
        r2, vx2, vy2 = lo.macroscopic(f, solid.astype(bool))
        assert np.abs(vx2 - vx).max() < tol and np.abs(vy2 - vy).max() < tol and np.abs(r2 - r).max() < tol
    assert np.abs(vel.cpu().numpy() - ref).max() < max(tol, 1e-6)
    assert np.abs(rho.cpu().numpy() - r).max() < max(tol, 1e-6)
    assert (vel.cpu().numpy()[solid.astype(bool)] == 0).all()


def test_lbm_generates_dataset_records(native, tmp_path):
    """boundary -> simulation domain (walls, 128 spare columns) -> flow -> cropped dataset pair -> TFRecords."""
    from fluid_generator import lbm
    import input.flow_input as fi
    from utils.synthetic import synth_boundary
    b = synth_boundary(2, 128, 256, seed=3)
    bb, s = lbm.simulate(b, steps=400)
    assert bb.shape == (2, 128, 256, 1) and tuple(s.shape) == (2, 128, 256, 2)
    s = s.cpu().numpy()
    assert np.isfinite(s).all() and 0.03 < s[..., 0].mean() < 0.12      # the channel carries the inlet's flux
    assert (s[bb[..., 0] == 1] == 0).all()                              # no flow inside solids
    assert (bb[:, 0] == 1).all() and (bb[:, -1] == 1).all()             # channel walls (:345-352)
    path = str(tmp_path / "lbm.tfrecords")
    fi.write_tfrecords(path, bb, s)
    got = list(fi.read_data(path))
    assert np.array_equal(got[1][0], bb[1]) and np.array_equal(got[1][1], s[1])


@pytest.mark.parametrize("B,H,W,C,N,ksize,stride,pro", [(2, 16, 32, 8, 16, 3, 1, 1), (3, 24, 40, 16, 32, 3, 1, 1), (2, 16, 32, 64, 128, 3, 1, 1),
                                                        (3, 8, 16, 128, 256, 3, 1, 1), (2, 32, 64, 8, 16, 3, 2, 1), (2, 16, 32, 64, 64, 1, 1, 1),
                                                        (40, 32, 64, 8, 8, 3, 1, 1), (2, 16, 32, 1, 8, 3, 1, 0)])
def test_bias_gradient_fused_into_wgrad(native, B, H, W, C, N, ksize, stride, pro):
    """fn_conv_wgrad_bias: db as one more accumulator of the tcgen05 wgrad kernel (A = a plane of ones) against the
    stand-alone bias kernel; dW unchanged by the extra accumulator."""
    g = torch.Generator().manual_seed(61)
    if C == 1:
        x = (torch.rand(B, H, W, C, generator=g) > 0.5).to("cuda", torch.bfloat16)
    else:
        x = torch.randn(B, H, W, C, generator=g).to("cuda", torch.bfloat16)
    OH, OW = -(-H // stride), -(-W // stride)
    dY = torch.randn(B, OH, OW, N, generator=g).to("cuda", torch.bfloat16)
    mult = 2 if pro in (1, 3) else 1
    dW0 = torch.empty((ksize * ksize, C * mult, N), device="cuda")
    dW1 = torch.empty_like(dW0)
    db = torch.full((N,), float("nan"), device="cuda")
    native.conv_wgrad(x, dY, dW0, ksize=ksize, stride=stride, prologue=pro)
    native.conv_wgrad(x, dY, dW1, ksize=ksize, stride=stride, prologue=pro, db=db)
    ref = torch.empty(N, device="cuda")
    native.bias_grad(dY, ref)
    torch.cuda.synchronize()
    assert native.tc_status() == 0
    assert torch.equal(dW0, dW1)
    assert rel_l2(db, ref) < 1e-5
    assert rel_l2(db, dY.double().sum(dim=(0, 1, 2))) < 1e-5


@pytest.mark.parametrize("npix_shape,F,Ca,kind,acc", [((3, 32, 64), 8, 8, 1, True), ((2, 16, 32), 16, 16, 1, False), ((2, 16, 32), 32, 32, 1, True),
                                                      ((2, 8, 16), 64, 64, 1, True), ((2, 16, 32), 16, 16, 3, True)])
def test_nin_backward_kernel_vs_two_step(native, npix_shape, F, Ca, kind, acc):
    """fn_nin_bwd (1x1 data gradient of the skip + adjoint of its concat prologue) vs conv(ksize=1) + fn_prologue_bwd."""
    g = torch.Generator().manual_seed(71)
    B, H, W = npix_shape
    dy = torch.randn(B, H, W, F, generator=g).to("cuda", torch.bfloat16)
    wn = ((torch.rand(2 * Ca, F, generator=g) * 2 - 1) * 0.3).to("cuda")
    a = torch.randn(B, H, W, Ca, generator=g).to("cuda", torch.bfloat16)
    base = torch.randn(B, H, W, Ca, generator=g).to("cuda", torch.bfloat16)
    fused = native.nin_bwd(dy, wn, a, kind, out=base.clone(), accumulate=acc)
    dA = (dy.double().reshape(-1, F) @ wn.double().t()).reshape(B, H, W, 2 * Ca).to(torch.bfloat16)
    two = native.prologue_bwd(dA, a, kind, out=base.clone(), accumulate=acc)
    torch.cuda.synchronize()
    assert rel_l2(fused, two) < 3e-3


@pytest.mark.parametrize("B,H,W", [(1, 36, 52), (2, 40, 72), (1, 23, 41), (1, 50, 100)])
def test_net_vs_oracle_on_sizes_that_are_not_multiples_of_16(native, B, H, W):
    """reference flow_architecture.py:136-142: when a level has an odd size the stride-2 conv pads (1,1), the transposed
    conv on the way up returns 2*ceil(n/2) rows, the skip `a` is zero-padded bottom/right to that size -- and the
    network's output is LARGER than its input (e.g. 36 -> 48 rows).  Shapes and values against the oracle."""
    x = fo.synth_boundary(B, H, W, seed=6)
    vs = fo.VariableStore(1234, torch.float64)
    ref = fo.inference(vs, torch.from_numpy(x).double(), 1.0)
    for impl in ("auto", "simt"):
        y = _net(native, x.astype(np.uint8), impl)
        assert tuple(y.shape) == tuple(ref.shape)
        assert rel_l2(y, ref) < NET_TOL, impl


@pytest.mark.parametrize("B,H,W", [(1, 48, 80), (3, 64, 96), (2, 80, 144)])
def test_net_vs_oracle_on_sizes_that_are_not_powers_of_two(native, B, H, W):
    """The network is fully convolutional for H, W multiples of 16 (four stride-2 levels); tiles at the right / bottom
    edges are ragged for every kernel family at these sizes."""
    x = fo.synth_boundary(B, H, W, seed=4)
    vs = fo.VariableStore(1234, torch.float64)
    ref = fo.inference(vs, torch.from_numpy(x).double(), 1.0)
    for impl in ("auto", "simt"):
        y = _net(native, x.astype(np.uint8), impl)
        assert tuple(y.shape) == (B, H, W, 2)
        assert rel_l2(y, ref) < NET_TOL, impl


def _pack_pass_major(native, w16, N, gated):
    half, quarter = w16.shape[1] // 2, w16.shape[1] // 4
    parts = []
    for ps in range(2):
        rows = torch.cat([w16[:, ps * quarter:(ps + 1) * quarter], w16[:, half + ps * quarter:half + (ps + 1) * quarter]], dim=1)
        parts.append(native.pack_conv_weights_tc(rows.contiguous(), None, 3, half, 0, N, gated))
    return torch.cat(parts)


@pytest.mark.parametrize("B,H,W,gate,Cres,pool", [(3, 8, 16, True, None, False), (2, 8, 16, False, None, False), (2, 16, 32, True, None, False),
                                                  (2, 8, 16, True, 128, True), (70, 8, 16, True, None, False), (1, 24, 40, False, None, False),
                                                  (80, 8, 16, True, None, False), (76, 8, 16, False, None, False)])
def test_two_pass_c256_kernel_vs_cuda_core_and_oracle(native, B, H, W, gate, Cres, pool):
    """conv_s1w.cu (C = 256: two K passes of 128 channels, N = 512 as two MMAs) against the fp64 oracle and the CUDA-core
    kernel -- the 8x16 level of the filter_size = 16 variant."""
    g = torch.Generator().manual_seed(81)
    C = 256
    N = 2 * C if gate else C
    x = bf16_round(torch.randn(B, H, W, C, generator=g, dtype=torch.float64))
    lim = (6.0 / (9 * (2 * C + N))) ** 0.5
    w = bf16_round((torch.rand(3, 3, 2 * C, N, generator=g, dtype=torch.float64) * 2 - 1) * lim * 2)
    b = (torch.rand(N, generator=g, dtype=torch.float64) * 2 - 1).float().double()
    res = None
    if gate:
        Cres = C if Cres is None else Cres
        RH, RW = (2 * H, 2 * W) if pool else (H, W)
        res = bf16_round(torch.randn(B, RH, RW, Cres, generator=g, dtype=torch.float64))
    w16 = w.reshape(9, 2 * C, N).to("cuda", torch.bfloat16).contiguous()
    w_tc = _pack_pass_major(native, w16, N, gate)
    ys = []
    for impl in (native.IMPL_TCGEN05, native.IMPL_SIMT):
        y = torch.full((B, H, W, C), float("nan"), device="cuda", dtype=torch.bfloat16)
        native.conv_fwd(x.to("cuda", torch.bfloat16), w16, b.to("cuda", torch.float32), y, B=B, H=H, W=W, Cin=C, Cout=N, ksize=3, stride=1,
                        prologue=1, epilogue=1 if gate else 0, impl=impl, w_tc=w_tc,
                        res=res.to("cuda", torch.bfloat16) if res is not None else None, res_pool=pool)
        torch.cuda.synchronize()
        assert native.tc_status() == 0
        assert native.last_used_tcgen05() == (impl == native.IMPL_TCGEN05)
        ys.append(y)
    assert rel_l2(ys[0], ys[1]) < 5e-3
    if B * H * W <= 1024:
        ref = oracle_fused_conv(x, w, b, 1, 1, 1 if gate else 0, None, None, res, pool)
        assert rel_l2(ys[0], ref) < LAYER_TOL


@pytest.mark.parametrize("B,H,W,SH,SW", [(2, 16, 32, 16, 32), (3, 8, 16, 8, 16), (1, 24, 48, 20, 40)])
def test_skip_pass_kernel_c128_vs_cuda_core_and_oracle(native, B, H, W, SH, SW):
    """conv_s1w.cu VAR_SKIP: conv_1 + nin skip at C = 128 (x pass + centre-tap skip pass), incl. a skip tensor smaller than
    the output (zero-padded bottom / right, flow_architecture.py:139-141)."""
    g = torch.Generator().manual_seed(91)
    C = 128
    x = bf16_round(torch.randn(B, H, W, C, generator=g, dtype=torch.float64))
    a = bf16_round(torch.randn(B, SH, SW, C, generator=g, dtype=torch.float64))
    lim = (6.0 / (9 * (2 * C + C))) ** 0.5
    w = bf16_round((torch.rand(3, 3, 2 * C, C, generator=g, dtype=torch.float64) * 2 - 1) * lim * 2)
    ws = bf16_round((torch.rand(2 * C, C, generator=g, dtype=torch.float64) * 2 - 1) * 0.1)
    b = (torch.rand(C, generator=g, dtype=torch.float64) * 2 - 1).float().double()
    w16 = w.reshape(9, 2 * C, C).to("cuda", torch.bfloat16).contiguous()
    ws16 = ws.to("cuda", torch.bfloat16).contiguous()
    w_tc = native.pack_conv_weights_tc(w16, ws16, 3, 2 * C, 2 * C, C, False)
    ys = []
    for impl in (native.IMPL_TCGEN05, native.IMPL_SIMT):
        y = torch.full((B, H, W, C), float("nan"), device="cuda", dtype=torch.bfloat16)
        native.conv_fwd(x.to("cuda", torch.bfloat16), w16, b.to("cuda", torch.float32), y, B=B, H=H, W=W, Cin=C, Cout=C, ksize=3, stride=1,
                        prologue=1, epilogue=0, impl=impl, w_tc=w_tc, skip=a.to("cuda", torch.bfloat16), w_skip=ws16)
        torch.cuda.synchronize()
        assert native.tc_status() == 0
        assert native.last_used_tcgen05() == (impl == native.IMPL_TCGEN05)
        ys.append(y)
    assert rel_l2(ys[0], ys[1]) < 5e-3
    ref = oracle_fused_conv(x, w, b, 1, 1, 0, a, ws)
    assert rel_l2(ys[0], ref) < LAYER_TOL


def test_compiled_training_step_matches_eager(native):
    """flow_net.CompiledTraining (forward + backward replayed from a CUDA graph, Adam outside) follows the eager step
    bit for bit over several steps (every reduction is ordered, the graph holds the same launches)."""
    import model.flow_architecture as arch
    import model.flow_net as flow_net

    def data(k):
        return flow_net.inputs(2, shape=(32, 64), seed=100 + k)

    try:
        arch.reset_variables(1234)
        flow_net.reset_training_state()
        flow_net.begin_training()
        eager_losses = []
        for k in range(3):
            b, s = data(k)
            p = flow_net.inference(b, 1.0)
            err = flow_net.loss_image_train(p, s)
            eager_losses.append(flow_net.train(err, 1e-3).item())
        want = {k: v.clone() for k, v in arch.global_variables().items()}
        flow_net.end_training()

        arch.reset_variables(1234)
        flow_net.reset_training_state()
        ct = flow_net.CompiledTraining(2, (32, 64))
        assert ct.launches_per_replay > 100
        losses = []
        for k in range(3):
            b, s = data(k)
            losses.append(ct.step(b, s, 1e-3).item())
        got = arch.global_variables()
        for a, b in zip(losses, eager_losses):           # the loss scalar is an atomicAdd over blocks: order-dependent rounding
            assert abs(a - b) <= 1e-6 * abs(b)
        for k in want:
            assert torch.equal(got[k], want[k]), k
    finally:
        flow_net.end_training()
        flow_net.reset_training_state()
        arch.reset_variables(1234)


def test_compiled_training_side_streams_change_nothing(native, monkeypatch):
    """The captured step issues the weight images and the parameter gradients on side streams of the graph
    (flow_architecture.prep_stream, BackwardPass.wstream).  At the full 128x256 size -- the persistent wgrad / dgrad kernels,
    several tiles per CTA -- the parameters after three steps must equal, bit for bit, those of the same graph captured
    on one stream (FLOWNET_TRAIN_STREAMS=0): a missing dependency or a recycled buffer would show."""
    import model.flow_architecture as arch
    import model.flow_net as flow_net

    def run(streams):
        monkeypatch.setenv("FLOWNET_TRAIN_STREAMS", streams)
        arch.reset_variables(1234)
        flow_net.reset_training_state()
        ct = flow_net.CompiledTraining(4, (128, 256))
        for k in range(3):
            b, s = flow_net.inputs(4, shape=(128, 256), seed=300 + k)
            ct.step(b, s, 1e-3)
        torch.cuda.synchronize()
        assert native.tc_status() == 0
        out = {k: v.clone() for k, v in arch.global_variables().items()}
        flow_net.end_training()
        flow_net.reset_training_state()
        return out

    try:
        want = run("0")
        for rep in range(2):
            got = run("1")
            for k in want:
                assert torch.equal(got[k], want[k]), (rep, k)
    finally:
        flow_net.end_training()
        flow_net.reset_training_state()
        arch.reset_variables(1234)


def test_compiled_training_with_dropout_draws_a_new_mask_per_replay(native):
    """keep_prob < 1 inside the captured step: the device-resident seed counter changes the mask on every replay (same
    inputs, different gradients), and the same counter value reproduces the same gradients."""
    import model.flow_architecture as arch
    import model.flow_net as flow_net
    try:
        arch.reset_variables(1234)
        flow_net.reset_training_state()
        ct = flow_net.CompiledTraining(2, (32, 64), keep_prob=0.7, dropout_seed=3)
        b, s = flow_net.inputs(2, shape=(32, 64), seed=5)
        flat = flow_net.ensure_training_state()
        grads = []
        for counter in (11, 12, 11):
            ct.static_boundary.copy_(b)
            ct.static_sflow.copy_(s)
            ct.seed_dev.fill_(counter)
            ct.graph.replay()
            torch.cuda.synchronize()
            grads.append(flat.grad.clone())
        assert native.tc_status() == 0
        assert torch.isfinite(grads[0]).all()
        assert torch.equal(grads[0], grads[2])
        assert rel_l2(grads[1], grads[0]) > 0.01          # another mask: measurably different gradients
        l0 = ct.step(b, s, 1e-3).item()
        assert np.isfinite(l0)
    finally:
        flow_net.end_training()
        flow_net.reset_training_state()
        arch.reset_variables(1234)


@pytest.mark.parametrize("B,gate,keep", [(5, True, 1.0), (5, False, 1.0), (64, True, 1.0), (3, True, 0.7), (4, "bwd", 1.0)])
def test_deep_level_channel_split_is_bit_identical(native, B, gate, keep):
    """conv_s1d NSPLIT = 2 (two CTAs per 8x16 tile, half of the output channels and of the weight stream each) against
    the unsplit kernel: same MMA order per output -> identical bits."""
    g = torch.Generator().manual_seed(51)
    C, H, W = 128, 8, 16
    gated = bool(gate)
    N = 2 * C if gated else C
    x = torch.randn(B, H, W, C, generator=g).to("cuda", torch.bfloat16)
    w16 = ((torch.rand(9, 2 * C, N, generator=g) * 2 - 1) * 0.03).to("cuda", torch.bfloat16)
    bias = (torch.rand(N, generator=g) - 0.5).to("cuda")
    res = torch.randn(B, H, W, C, generator=g).to("cuda", torch.bfloat16) if gated else None
    w_tc = native.pack_conv_weights_tc(w16, None, 3, 2 * C, 0, N, gated)
    epi = native.EPI_GATE_BWD if gate == "bwd" else (native.EPI_GATE_RES if gated else native.EPI_BIAS)
    cy = N if gate == "bwd" else C
    ys = []
    try:
        for split in (1, 0):
            native.lib.fn_debug_set_nsplit(split)
            y = torch.full((B, H, W, cy), float("nan"), device="cuda", dtype=torch.bfloat16)
            native.conv_fwd(x, w16, bias, y, B=B, H=H, W=W, Cin=C, Cout=N, ksize=3, stride=1, prologue=1, epilogue=epi,
                            impl=native.IMPL_TCGEN05, w_tc=w_tc, res=res, keep_prob=keep, dropout_seed=2)
            torch.cuda.synchronize()
            assert native.tc_status() == 0
            ys.append(y)
    finally:
        native.lib.fn_debug_set_nsplit(1)
    assert not torch.isnan(ys[0].float()).any()
    assert torch.equal(ys[0], ys[1])
