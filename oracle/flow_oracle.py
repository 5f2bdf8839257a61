"""CPU oracle for the flow_net hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product path (steady-state-flow-with-neural-nets_b200/)
never does, and has no CPU fallback.

What it is: a restatement, on torch-CPU tensors (fp64 for parity, fp32 for timing), of the
TensorFlow-1.x graph the reference builds in

    /root/reference/model/flow_architecture.py   (concat_elu :14-17, conv_layer :57-68,
        transpose_conv_layer :70-87, fc_layer :89-104, nin :106-111, res_block :129-165,
        conv_res :167-248)
    /root/reference/model/flow_net.py            (inference :33-37, loss_image :39-42,
        train :44-46)

The arithmetic of the reference lives in TensorFlow (third party, un-vendored, version
unpinned; API usage implies 1.5 <= TF <= 1.15).  TensorFlow is not installable here and
the reference ships no golden vectors, checkpoints or asserting tests, so:

    **PARITY UNPINNED at the TensorFlow boundary.**

The TF op semantics restated here (SAME padding asymmetry, conv2d_transpose phase taps,
weight layouts, front channel pad, Xavier limits incl. biases, elu, TF1 dropout scaling,
l2_loss, TF1 ApplyAdam) are pinned by SURVEY.md App. B and cross-checked against an
independent naive-loop implementation (oracle/naive_numpy.py) and adjointness tests.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- helpers
def int_shape(x):
    """flow_architecture.py:11-12"""
    return [int(s) for s in x.shape]


def _same_pads(n_in: int, k: int, s: int):
    """TF 'SAME' padding: out = ceil(in/s); total = max((out-1)s+k-in,0);
    before = total//2; after = total-before."""
    out = -(-n_in // s)
    total = max((out - 1) * s + k - n_in, 0)
    before = total // 2
    return before, total - before, out


def _nhwc_to_nchw(x):
    return x.permute(0, 3, 1, 2)


def _nchw_to_nhwc(x):
    return x.permute(0, 2, 3, 1)


# --------------------------------------------------------------------------- variables
class VariableStore:
    """Name-keyed stand-in for the TF graph's variable collection.

    Variables are created lazily, in call order, exactly as tf.get_variable would
    (flow_architecture.py:44-55): '<scope>/weights' then '<scope>/biases'.  The initializer
    is tf.contrib.layers.xavier_initializer[_conv2d]() == variance_scaling(factor=1,
    mode=FAN_AVG, uniform=True) for weights AND biases (flow_architecture.py:61-62,74-75,
    99-100): limit = sqrt(6/(fan_in+fan_out)); for a 1-D shape [n], fan_in = fan_out = n.
    """

    def __init__(self, seed: int = 1234, dtype=torch.float64):
        self.rng = np.random.default_rng(seed)
        self.dtype = dtype
        self.vars: "OrderedDict[str, torch.Tensor]" = OrderedDict()

    @staticmethod
    def xavier_limit(shape):
        shape = list(shape)
        if len(shape) == 1:
            fan_in = fan_out = shape[0]
        else:
            rf = 1
            for d in shape[:-2]:
                rf *= d
            fan_in = shape[-2] * rf
            fan_out = shape[-1] * rf
        return math.sqrt(6.0 / (fan_in + fan_out))

    def get(self, name, shape):
        if name in self.vars:
            v = self.vars[name]
            assert list(v.shape) == list(shape), (name, v.shape, shape)
            return v
        lim = self.xavier_limit(shape)
        # drawn as fp32 (TF variables are float32), then widened for the fp64 oracle
        arr = self.rng.uniform(-lim, lim, size=shape).astype(np.float32)
        v = torch.from_numpy(arr).to(self.dtype)
        self.vars[name] = v
        return v

    def num_params(self):
        return sum(v.numel() for v in self.vars.values())

    def astype(self, dtype):
        out = VariableStore.__new__(VariableStore)
        out.rng = None
        out.dtype = dtype
        out.vars = OrderedDict((k, v.to(dtype)) for k, v in self.vars.items())
        return out


# --------------------------------------------------------------------------- layers
def elu(x):
    return torch.where(x > 0, x, torch.expm1(torch.clamp(x, max=0)))


def concat_elu(x):
    """flow_architecture.py:14-17 : elu(concat([x,-x], channel axis))"""
    return elu(torch.cat([x, -x], dim=-1))


def concat_relu(x):
    """tf.nn.crelu"""
    return torch.relu(torch.cat([x, -x], dim=-1))


def set_nonlinearity(name):
    """flow_architecture.py:19-29"""
    if name == "concat_elu":
        return concat_elu
    if name == "elu":
        return elu
    if name == "concat_relu":
        return concat_relu
    if name == "relu":
        return torch.relu
    raise ValueError("nonlinearity " + name + " is not supported")


def conv2d_same(x, w, stride):
    """tf.nn.conv2d(x NHWC, w HWIO, strides=[1,s,s,1], 'SAME') -- cross-correlation."""
    k = w.shape[0]
    pt, pb, _ = _same_pads(x.shape[1], k, stride)
    pl, pr, _ = _same_pads(x.shape[2], k, stride)
    xn = F.pad(_nhwc_to_nchw(x), (pl, pr, pt, pb))
    y = F.conv2d(xn, w.permute(3, 2, 0, 1), stride=stride)
    return _nchw_to_nhwc(y)


def conv2d_transpose_same(x, w, stride):
    """tf.nn.conv2d_transpose(x, w [k,k,Cout,Cin], out=[B,sH,sW,Cout], 'SAME').

    It is the adjoint of conv2d_same acting on a (sH,sW) image; for even
    sH and k=3,s=2 that forward conv pads (0,1), so out[I] += in[i]*W[di] for I = s*i+di-pad_before.
    """
    k = w.shape[0]
    h, wd = x.shape[1], x.shape[2]
    H, W = h * stride, wd * stride
    pt, _, oh = _same_pads(H, k, stride)
    pl, _, ow = _same_pads(W, k, stride)
    assert oh == h and ow == wd
    full = F.conv_transpose2d(_nhwc_to_nchw(x), w.permute(3, 2, 0, 1), stride=stride)
    y = full[:, :, pt:pt + H, pl:pl + W]
    # conv_transpose2d's natural output is (h-1)*s+k; SAME may need more rows than that
    if y.shape[2] < H or y.shape[3] < W:
        y = F.pad(y, (0, W - y.shape[3], 0, H - y.shape[2]))
    return _nchw_to_nhwc(y)


def conv_layer(vs, inputs, kernel_size, stride, num_features, idx, nonlinearity=None):
    """flow_architecture.py:57-68"""
    cin = inputs.shape[3]
    w = vs.get(f"{idx}_conv/weights", [kernel_size, kernel_size, cin, num_features])
    b = vs.get(f"{idx}_conv/biases", [num_features])
    y = conv2d_same(inputs, w, stride) + b
    if nonlinearity is not None:
        y = nonlinearity(y)
    return y


def transpose_conv_layer(vs, inputs, kernel_size, stride, num_features, idx, nonlinearity=None):
    """flow_architecture.py:70-87"""
    cin = inputs.shape[3]
    w = vs.get(f"{idx}_trans_conv/weights", [kernel_size, kernel_size, num_features, cin])
    b = vs.get(f"{idx}_trans_conv/biases", [num_features])
    y = conv2d_transpose_same(inputs, w, stride) + b
    if nonlinearity is not None:
        y = nonlinearity(y)
    return y


def fc_layer(vs, inputs, hiddens, idx, nonlinearity=None, flat=False):
    """flow_architecture.py:89-104 (flat=True / nonlinearity branches are unused upstream)"""
    if flat:
        inputs = inputs.reshape(inputs.shape[0], -1)
    dim = inputs.shape[1]
    w = vs.get(f"{idx}_fc/weights", [dim, hiddens])
    b = vs.get(f"{idx}_fc/biases", [hiddens])
    y = inputs @ w + b
    if nonlinearity is not None:
        y = nonlinearity(y)
    return y


def nin(vs, x, num_units, idx):
    """flow_architecture.py:106-111 : 1x1 conv as a matmul over flattened pixels"""
    s = list(x.shape)
    y = fc_layer(vs, x.reshape(-1, s[-1]), num_units, idx)
    return y.reshape(s[:-1] + [num_units])


def avg_pool_2x2_same(x):
    """tf.nn.avg_pool(x,[1,2,2,1],[1,2,2,1],'SAME'): edge windows divide by the valid count."""
    y = F.avg_pool2d(_nhwc_to_nchw(x), 2, 2, ceil_mode=True, count_include_pad=False)
    return _nchw_to_nhwc(y)


def dropout_tf1(x, keep_prob, gen):
    """tf.nn.dropout (TF1): x / keep_prob * floor(keep_prob + U[0,1))"""
    u = torch.rand(x.shape, generator=gen, dtype=x.dtype)
    return x / keep_prob * torch.floor(keep_prob + u)


def res_block(vs, x, a=None, filter_size=16, nonlinearity=concat_elu, keep_p=1.0, stride=1,
              gated=False, name="resnet", gen=None, mask_seed=None):
    """flow_architecture.py:129-165.  mask_seed: use the CUDA path's counter-based dropout mask with this seed
    (oracle/dropout_oracle.py) instead of a torch RNG draw, so both sides drop the same elements."""
    orig_x = x
    if x.shape[3] == 1:
        x_1 = conv_layer(vs, x, 3, stride, filter_size, name + "_conv_1")
    else:
        x_1 = conv_layer(vs, nonlinearity(x), 3, stride, filter_size, name + "_conv_1")
    if a is not None:
        dh = x_1.shape[1] - a.shape[1]
        dw = x_1.shape[2] - a.shape[2]
        assert dh >= 0 and dw >= 0
        a = F.pad(a, (0, 0, 0, dw, 0, dh))
        x_1 = x_1 + nin(vs, nonlinearity(a), filter_size, name + "_nin")
    x_1 = nonlinearity(x_1)
    if keep_p < 1.0:
        if mask_seed is not None:
            from oracle import dropout_oracle
            x_1 = x_1 * torch.from_numpy(dropout_oracle.dropout_scale(mask_seed, tuple(x_1.shape), keep_p)).to(x_1.dtype)
        else:
            x_1 = dropout_tf1(x_1, keep_p, gen)
    if not gated:
        x_2 = conv_layer(vs, x_1, 3, 1, filter_size, name + "_conv_2")
    else:
        x_2 = conv_layer(vs, x_1, 3, 1, filter_size * 2, name + "_conv_2")
        x_2_1, x_2_2 = torch.split(x_2, filter_size, dim=3)
        x_2 = x_2_1 * torch.sigmoid(x_2_2)
    if orig_x.shape[2] > x_2.shape[2]:
        orig_x = avg_pool_2x2_same(orig_x)
    out_filter = filter_size
    in_filter = orig_x.shape[3]
    if out_filter != in_filter:
        orig_x = F.pad(orig_x, (out_filter - in_filter, 0))   # channels padded IN FRONT
    return orig_x + x_2


def conv_res(vs, inputs, nr_res_blocks=1, keep_prob=1.0, nonlinearity_name="concat_elu",
             gated=True, filter_size=8, gen=None, taps=None, dropout_seed=None):
    """flow_architecture.py:167-248.  `filter_size` (hard-coded 8 upstream, :174) is exposed
    for BASELINE config 5.  `taps`, if a dict, receives named intermediate activations."""
    nonlinearity = set_nonlinearity(nonlinearity_name)
    kw = dict(nonlinearity=nonlinearity, keep_p=keep_prob, gated=gated, gen=gen)
    nblock = [0]

    def block(vs, x, **k):                         # numbers the blocks: block n drops with seed dropout_seed + n
        nblock[0] += 1
        return res_block(vs, x, mask_seed=None if dropout_seed is None else dropout_seed + nblock[0], **k)

    def tap(name, t):
        if taps is not None:
            taps[name] = t
        return t

    a = []
    x = inputs
    for i in range(nr_res_blocks):
        x = tap(f"resnet_1_{i}", block(vs, x, filter_size=filter_size, name=f"resnet_1_{i}", **kw))
    for lvl in (2, 3, 4, 5):
        a.append(x)
        filter_size = 2 * filter_size
        x = tap(f"resnet_{lvl}_downsample",
                block(vs, x, filter_size=filter_size, stride=2, name=f"resnet_{lvl}_downsample", **kw))
        for i in range(nr_res_blocks):
            x = tap(f"resnet_{lvl}_{i}",
                    block(vs, x, filter_size=filter_size, name=f"resnet_{lvl}_{i}", **kw))
    for k in (1, 2, 3, 4):
        filter_size = int(filter_size / 2)
        x = tap(f"up_conv_{k}", transpose_conv_layer(vs, x, 3, 2, filter_size, f"up_conv_{k}"))
        for i in range(nr_res_blocks):
            x = tap(f"resnet_up_{k}_{i}",
                    block(vs, x, a=a[-k] if i == 0 else None, filter_size=filter_size,
                              name=f"resnet_up_{k}_{i}", **kw))
    x = conv_layer(vs, x, 3, 1, 2, "last_conv")
    return torch.tanh(x)


def inference(vs, boundary, keep_prob, nr_res_blocks=1, nonlinearity="concat_elu",
              gated_res=True, filter_size=8, gen=None, taps=None, dropout_seed=None):
    """flow_net.py:33-37 with FLAGS.model == 'res' and the model flags as keywords."""
    return conv_res(vs, boundary, nr_res_blocks=nr_res_blocks, keep_prob=keep_prob,
                    nonlinearity_name=nonlinearity, gated=gated_res, filter_size=filter_size,
                    gen=gen, taps=taps, dropout_seed=dropout_seed)


def loss_image(sflow_p, sflow):
    """flow_net.py:39-42 : tf.nn.l2_loss = sum(d^2)/2, no mean."""
    d = sflow_p - sflow
    return (d * d).sum() / 2


def adam_step_tf1(params, grads, m, v, t, lr=1e-4, beta1=0.9, beta2=0.999, eps=1e-8):
    """TF1 ApplyAdam (flow_net.py:45), in place; t is the 1-based step."""
    lr_t = lr * math.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    for k in params:
        g = grads[k]
        m[k].mul_(beta1).add_(g, alpha=1 - beta1)
        v[k].mul_(beta2).addcmul_(g, g, value=1 - beta2)
        params[k].sub_(lr_t * m[k] / (v[k].sqrt() + eps))


# --------------------------------------------------------------------------- inputs
def synth_boundary(B, H, W, seed=0):
    """Synthetic binary boundary images in the style of the reference's LBM generator
    (/root/reference/fluid_generator/generate_xiao_fluid_flow.cpp:268-352):
    top/bottom channel walls, 4*(W/256)*(H/128) random ellipses / rectangles per image,
    half-sizes in [20,40)*min(H/128,1), none left of x=32 or above y=32 (scaled)."""
    out = np.zeros((B, H, W, 1), dtype=np.float32)
    sc = min(H / 128.0, 1.0)
    yy, xx = np.mgrid[0:H, 0:W]
    for b in range(B):
        rng = np.random.default_rng(seed + b)
        g = np.zeros((H, W), dtype=bool)
        g[0, :] = True
        g[H - 1, :] = True
        nobj = max(1, int(round(4 * (W / 256.0) * (H / 128.0))))
        for _ in range(nobj):
            kind = rng.integers(0, 2)
            rx = rng.uniform(20, 40) * sc
            ry = rng.uniform(20, 40) * sc
            lo_x, hi_x = 32 * sc + rx, max(W - rx, 32 * sc + rx + 1)
            lo_y, hi_y = 32 * sc, max(H - ry / 2, 32 * sc + 1)
            cx = rng.uniform(lo_x, hi_x)
            cy = rng.uniform(lo_y, hi_y)
            if kind == 0:
                g |= ((xx - cx) / rx) ** 2 + ((yy - cy) / ry) ** 2 <= 1.0
            else:
                g |= (np.abs(xx - cx) <= rx) & (np.abs(yy - cy) <= ry)
        out[b, :, :, 0] = g
    return out


def car_boundary_from_txt(path, H=128, W=256):
    """cars/car_NNN.txt (header '300 100', then 300 x-major rows of 100) placed as the LBM
    generator does (generate_xiao_fluid_flow.cpp:250-262, :348-352) and cropped as
    utils/flow_reader.py:12-17 does: g[28+j, 32+x] = txt[x][j]; walls at rows 0, 127."""
    with open(path) as f:
        toks = f.read().split()
    nx, ny = int(toks[0]), int(toks[1])
    vals = np.array(toks[2:2 + nx * ny], dtype=np.float32).reshape(nx, ny)
    g = np.zeros((128, 384), dtype=np.float32)
    g[28:28 + ny, 32:32 + nx] = vals.T
    g[0, :] = 1
    g[127, :] = 1
    return g[:H, :W]
