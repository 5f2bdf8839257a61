"""Second, independent restatement of the flow_net forward in plain NumPy -- TEST INFRASTRUCTURE.

Written straight from the index formulas in SURVEY.md App. B (no library convolution, no
shared code with oracle/flow_oracle.py) so the two can be cross-checked in fp64.  The
per-pixel ops are explicit Python loops over output pixels and taps; use small shapes.

Follows /root/reference/model/flow_architecture.py:14-17, 57-111, 129-248.
"""
from __future__ import annotations

import math

import numpy as np


def same_pad(n, k, s):
    out = (n + s - 1) // s
    tot = max((out - 1) * s + k - n, 0)
    return tot // 2, out


def conv2d_same(x, w, b, stride):
    """y[n,i,j,o] = sum_{di,dj,c} x[n, s*i+di-pb, s*j+dj-pb, c] * w[di,dj,c,o] + b[o]"""
    B, H, W, C = x.shape
    k = w.shape[0]
    pt, OH = same_pad(H, k, stride)
    pl, OW = same_pad(W, k, stride)
    y = np.zeros((B, OH, OW, w.shape[3]), dtype=x.dtype)
    for i in range(OH):
        for j in range(OW):
            acc = np.zeros((B, w.shape[3]), dtype=x.dtype)
            for di in range(k):
                ii = stride * i + di - pt
                if ii < 0 or ii >= H:
                    continue
                for dj in range(k):
                    jj = stride * j + dj - pl
                    if jj < 0 or jj >= W:
                        continue
                    acc += x[:, ii, jj, :] @ w[di, dj]
            y[:, i, j, :] = acc + b
    return y


def conv2d_transpose_same(x, w, b, stride):
    """Scatter form of App. B.2: out[n, s*i+di-pb, s*j+dj-pb, o] += x[n,i,j,c] * w[di,dj,o,c]"""
    B, h, wd, C = x.shape
    k = w.shape[0]
    H, W = h * stride, wd * stride
    pt, _ = same_pad(H, k, stride)
    pl, _ = same_pad(W, k, stride)
    y = np.zeros((B, H, W, w.shape[2]), dtype=x.dtype)
    for i in range(h):
        for j in range(wd):
            for di in range(k):
                I = stride * i + di - pt
                if I < 0 or I >= H:
                    continue
                for dj in range(k):
                    J = stride * j + dj - pl
                    if J < 0 or J >= W:
                        continue
                    y[:, I, J, :] += x[:, i, j, :] @ w[di, dj].T
    return y + b


def elu(x):
    return np.where(x > 0, x, np.expm1(np.minimum(x, 0)))


def concat_elu(x):
    return elu(np.concatenate([x, -x], axis=-1))


def avg_pool_2x2_same(x):
    B, H, W, C = x.shape
    OH, OW = (H + 1) // 2, (W + 1) // 2
    y = np.zeros((B, OH, OW, C), dtype=x.dtype)
    for i in range(OH):
        for j in range(OW):
            win = x[:, 2 * i:min(2 * i + 2, H), 2 * j:min(2 * j + 2, W), :]
            y[:, i, j, :] = win.mean(axis=(1, 2))
    return y


class Vars:
    """Pulls variables by name from a dict of numpy arrays (shapes as the TF graph has them)."""

    def __init__(self, d):
        self.d = d

    def __call__(self, name):
        return self.d[name]


def res_block(v, x, a, F, stride, name):
    orig = x
    inp = x if x.shape[3] == 1 else concat_elu(x)
    x1 = conv2d_same(inp, v(name + "_conv_1_conv/weights"), v(name + "_conv_1_conv/biases"), stride)
    if a is not None:
        ap = np.zeros(x1.shape[:3] + (a.shape[3],), dtype=x.dtype)
        ap[:, :a.shape[1], :a.shape[2], :] = a
        x1 = x1 + concat_elu(ap) @ v(name + "_nin_fc/weights") + v(name + "_nin_fc/biases")
    x1 = concat_elu(x1)
    x2 = conv2d_same(x1, v(name + "_conv_2_conv/weights"), v(name + "_conv_2_conv/biases"), 1)
    x2 = x2[..., :F] * (1.0 / (1.0 + np.exp(-x2[..., F:])))
    if orig.shape[2] > x2.shape[2]:
        orig = avg_pool_2x2_same(orig)
    cin = orig.shape[3]
    if cin != F:
        pad = np.zeros(orig.shape[:3] + (F - cin,), dtype=x.dtype)
        orig = np.concatenate([pad, orig], axis=-1)
    return orig + x2


def conv_res(vars_dict, inputs, nr_res_blocks=1, filter_size=8):
    """Default-flag forward (concat_elu, gated, keep_prob=1)."""
    v = Vars(vars_dict)
    a = []
    x = inputs
    F = filter_size
    for i in range(nr_res_blocks):
        x = res_block(v, x, None, F, 1, f"resnet_1_{i}")
    for lvl in (2, 3, 4, 5):
        a.append(x)
        F *= 2
        x = res_block(v, x, None, F, 2, f"resnet_{lvl}_downsample")
        for i in range(nr_res_blocks):
            x = res_block(v, x, None, F, 1, f"resnet_{lvl}_{i}")
    for k in (1, 2, 3, 4):
        F //= 2
        x = conv2d_transpose_same(x, v(f"up_conv_{k}_trans_conv/weights"),
                                  v(f"up_conv_{k}_trans_conv/biases"), 2)
        for i in range(nr_res_blocks):
            x = res_block(v, x, a[-k] if i == 0 else None, F, 1, f"resnet_up_{k}_{i}")
    x = conv2d_same(x, v("last_conv_conv/weights"), v("last_conv_conv/biases"), 1)
    return np.tanh(x)


def xavier_limit(shape):
    if len(shape) == 1:
        fi = fo = shape[0]
    else:
        rf = int(np.prod(shape[:-2]))
        fi, fo = shape[-2] * rf, shape[-1] * rf
    return math.sqrt(6.0 / (fi + fo))
