"""functions used to construct different architectures -- B200-native.

Mirrors /root/reference/model/flow_architecture.py name for name (int_shape, concat_elu,
set_nonlinearity, _variable, conv_layer, transpose_conv_layer, fc_layer, nin, res_block, conv_res)
but executes eagerly on CUDA tensors: every function launches hand-written sm_100a kernels through
the C ABI of libflownet.so (include/flownet.h).  Activations are NHWC; they are stored as bf16
between kernels and accumulated in fp32; the network output is fp32.

Fusion happens where the reference composes ops: res_block() (reference :129-165) issues exactly
two kernels -- conv_1 with the nonlinearity as A-operand prologue and the nin skip folded in as
extra GEMM-K, then conv_2 with nonlinearity prologue and gate*sigmoid + shortcut epilogue.  The
stand-alone layer functions remain available and compute the same values un-fused.
"""
from __future__ import annotations

import math
import os
from collections import OrderedDict

import numpy as np
import torch

import flownet_native as native
from model.flags import FLAGS  # noqa: F401  (the reference exposes FLAGS here too, :9)

# --------------------------------------------------------------------------- configuration
_IMPL_NAMES = {"auto": native.IMPL_AUTO, "simt": native.IMPL_SIMT, "tc": native.IMPL_TCGEN05}
IMPL = _IMPL_NAMES[os.environ.get("FLOWNET_IMPL", "auto")]


# when a list, every conv / tconv launch appends a record with CUDA events around it (bench.py's
# per-kernel roofline pass); None in normal operation
_PROFILE = None


# measured execution time of one tcgen05.mma (M = 128, K = 16, bf16) by accumulator width N, in SM clocks
# (tools/mma_bench.cu, profiles/r2_mma_bench3.log): the A operand's 4 KB shared-memory read bounds the narrow shapes
_MMA_CLK = {16: 39.0, 32: 42.0, 64: 48.0, 128: 64.0, 256: 128.0}


def _mma_clk(blocks, ksteps, n):
    """MMA-unit clocks of `blocks` M=128 tiles x `ksteps` K=16 steps at accumulator width n (rounded up to a measured
    width; n > 256 = several instructions) -- the numerator of bench.py's MMA-rate floor."""
    n16 = -(-int(n) // 16) * 16
    clk, rem = 0.0, n16
    while rem > 0:
        w = min(rem, 256)
        clk += _MMA_CLK[min(k for k in _MMA_CLK if k >= w)]
        rem -= w
    return float(blocks) * float(ksteps) * clk


def _profiled(record, launch):
    if _PROFILE is None:
        launch()
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    launch()
    e1.record()
    record["events"] = (e0, e1)
    record["tcgen05"] = native.last_used_tcgen05()
    _PROFILE.append(record)


# when a list, res_block / transpose_conv_layer / conv_res record what the backward pass needs
# (model/flow_backward.py); None in inference
TAPE = None


def set_impl(name: str):
    """'auto' (tcgen05 wherever a kernel exists), 'simt' (CUDA-core cross-check), 'tc' (fail if no tcgen05 kernel)."""
    global IMPL
    IMPL = _IMPL_NAMES[name]


def int_shape(x):
    """reference :11-12"""
    return list(map(int, x.shape))


# --------------------------------------------------------------------------- variables
class _VariableStore:
    """The TF graph's variable collection: name -> fp32 CUDA tensor, created lazily in call order
    with the reference's initializer (xavier for weights AND biases, reference :61-62,:74-75,:99-100;
    SURVEY App. B.3).  Names, shapes and layouts are TF's, so a checkpoint dump loads 1:1."""

    def __init__(self, seed=1234):
        self.reset(seed)

    def reset(self, seed=1234):
        self.seed = seed
        self.rng = np.random.default_rng(seed)
        self.vars = OrderedDict()
        self.version = 0
        self._derived = {}
        self._preload = {}

    def preload(self, values):
        """Initial values (name -> numpy fp32) for variables the lazy graph has not created yet (checkpoint restore)."""
        self._preload.update(values)

    def get(self, name, shape, device):
        v = self.vars.get(name)
        if v is not None:
            if list(v.shape) != list(shape):
                raise ValueError(f"variable {name} has shape {list(v.shape)}, requested {list(shape)}")
            return v
        lim = xavier_limit(shape)
        arr = self.rng.uniform(-lim, lim, size=shape).astype(np.float32)     # always drawn: keeps later draws in step
        pre = self._preload.pop(name, None)
        if pre is not None:
            if list(pre.shape) != list(shape):
                raise ValueError(f"checkpoint value for {name} has shape {list(pre.shape)}, graph wants {list(shape)}")
            arr = np.ascontiguousarray(pre, dtype=np.float32)
        v = torch.from_numpy(arr).to(device)
        self.vars[name] = v
        return v

    def touch(self):
        """Call after mutating variables in place (optimizer step, checkpoint load)."""
        self.version += 1
        self._derived.clear()

    def derived(self, key, make):
        ent = self._derived.get(key)
        if ent is None:
            ent = weight_prep(make)
            self._derived[key] = ent
        return ent


# ---- weight preparation on a side stream of the training graphs
# The bf16 / tensor-core images of the variables depend on nothing but the variables, yet traced in place they sit
# between the convs as ~120 dependent 3-us launches per training step (ncu launch list, profiles/r2_launches_train.csv).
# While a training graph is being captured they are issued on a side stream forked at the graph's start: in the graph
# they form a chain of their own that runs ahead of (and beside) the conv chain; each consumer waits only for its image.
_PREP = {"stream": None, "keep": None}


def weight_prep(make):
    """Run make() (casts, packs: anything that reads variables only) on the capture's side stream when there is one."""
    side = _PREP["stream"]
    if side is None:
        return make()
    with torch.cuda.stream(side):
        ent = make()
        ev = torch.cuda.Event()
        ev.record(side)
    torch.cuda.current_stream().wait_event(ev)
    # the side stream's allocator must not hand these blocks to a later image while a conv on the main chain reads them
    _PREP["keep"].append(ent)
    return ent


class prep_stream:
    """with prep_stream(keep): ... inside a torch.cuda.graph capture -- forks the weight-preparation stream at entry and
    joins it at exit; `keep` (a list owned by the graph's holder) receives every tensor made on it."""

    def __init__(self, keep):
        self.keep = keep

    def __enter__(self):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        _PREP["stream"], _PREP["keep"] = side, self.keep
        return self

    def __exit__(self, *exc):
        side = _PREP["stream"]
        _PREP["stream"] = _PREP["keep"] = None
        torch.cuda.current_stream().wait_stream(side)
        return False


def xavier_limit(shape):
    shape = list(shape)
    if len(shape) == 1:
        fan_in = fan_out = shape[0]
    else:
        rf = 1
        for d in shape[:-2]:
            rf *= d
        fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
    return math.sqrt(6.0 / (fan_in + fan_out))


VARIABLES = _VariableStore()


def reset_variables(seed=1234):
    """Start a fresh 'graph': forget all variables and re-seed the initializer."""
    VARIABLES.reset(seed)


def global_variables():
    return VARIABLES.vars


def _variable(name, shape, initializer=None, device="cuda"):
    """reference :44-55 (tf.get_variable).  `initializer` is accepted for signature parity; the
    reference only ever passes the Xavier initializer, which is what is used."""
    return VARIABLES.get(name, shape, device)


# --------------------------------------------------------------------------- nonlinearities
def concat_elu(x):
    """like concatenated ReLU (http://arxiv.org/abs/1603.05201), but then with ELU -- reference :14-17"""
    return native.nonlinearity(native.to_bf16(x), native.PRO_CONCAT_ELU)


def elu(x):
    return native.nonlinearity(native.to_bf16(x), native.PRO_ELU)


def concat_relu(x):
    return native.nonlinearity(native.to_bf16(x), native.PRO_CONCAT_RELU)


def relu(x):
    return native.nonlinearity(native.to_bf16(x), native.PRO_RELU)


_PROLOGUE_OF = {concat_elu: native.PRO_CONCAT_ELU, elu: native.PRO_ELU,
                concat_relu: native.PRO_CONCAT_RELU, relu: native.PRO_RELU}


def set_nonlinearity(name):
    """reference :19-29"""
    if name == 'concat_elu':
        return concat_elu
    elif name == 'elu':
        return elu
    elif name == 'concat_relu':
        return concat_relu
    elif name == 'relu':
        return relu
    else:
        raise ValueError('nonlinearity ' + name + ' is not supported')


def _pro_mult(pro):
    return 2 if pro in (native.PRO_CONCAT_ELU, native.PRO_CONCAT_RELU) else 1


# --------------------------------------------------------------------------- fused conv launcher
def _conv_weights(scope, ksize, cin_eff, cout, device, skip_scope=None, kskip_eff=0, gated=False, pass_major=False):
    """bf16 GEMM-B images of a conv's variables (+ the nin it absorbs), cached per variable version.
    pass_major: the 512-row (C = 256, concat) weights of the two-pass kernel conv_s1w.cu -- the tensor-core image is
    that of the K rows {0..127, 256..383} followed by that of {128..255, 384..511}."""
    w = _variable(scope + '/weights', [ksize, ksize, cin_eff, cout], device=device)
    b = _variable(scope + '/biases', [cout], device=device)
    ws = bs = None
    if skip_scope is not None:
        ws = _variable(skip_scope + '/weights', [kskip_eff, cout], device=device)
        bs = _variable(skip_scope + '/biases', [cout], device=device)

    def make():
        w16 = w.reshape(ksize * ksize, cin_eff, cout).to(torch.bfloat16).contiguous()
        ws16 = ws.to(torch.bfloat16).contiguous() if ws is not None else None
        bias = (b + bs).contiguous() if bs is not None else b.contiguous()
        w_tc = None
        if pass_major:
            half, quarter = cin_eff // 2, cin_eff // 4
            parts = []
            for ps in range(2):
                rows = torch.cat([w16[:, ps * quarter:(ps + 1) * quarter], w16[:, half + ps * quarter:half + (ps + 1) * quarter]], dim=1)
                parts.append(native.pack_conv_weights_tc(rows.contiguous(), None, ksize, half, 0, cout, gated))
            w_tc = torch.cat(parts)
        elif cin_eff % 16 == 0 and kskip_eff % 16 == 0:
            w_tc = native.pack_conv_weights_tc(w16, ws16, ksize, cin_eff, kskip_eff, cout, gated)
        return w16, ws16, bias, w_tc

    return VARIABLES.derived(("conv", scope, skip_scope, bool(pass_major), bool(gated)), make)


def _fused_conv(x, scope, ksize, stride, cout, prologue, epilogue, skip=None, skip_scope=None, res=None,
                res_pool=False, keep_prob=1.0, dropout_seed=0, query=False):
    """One fused launch.  query=True returns whether it would run on the tensor cores instead of launching."""
    # uint8 stays uint8 only for the one launch that reads it directly, the 1 -> 8 channel head conv (fn_conv_desc.x_u8)
    if not (x.dtype == torch.uint8 and x.shape[3] == 1 and cout == 8 and ksize == 3 and stride == 1 and prologue == native.PRO_NONE
            and epilogue == native.EPI_BIAS and skip is None and keep_prob >= 1.0):
        x = native.to_bf16(x)
    B, H, W, cin = int_shape(x)
    mult = _pro_mult(prologue)
    kskip = 0
    if skip is not None:
        skip = native.to_bf16(skip)
        kskip = skip.shape[3] * mult
    gated = epilogue in (native.EPI_GATE_RES, native.EPI_GATE_BWD)
    # the two-pass 256-channel kernel reads its weights pass-major: the library says whether it will take this launch
    # (fn_conv_weights_layout -- the dispatch's own acceptance test; placeholders stand in for the tensors not made yet)
    pass_major = False
    if cin * mult == 512 and skip is None and IMPL != native.IMPL_SIMT:
        pass_major = bool(native.conv_fwd(x, x, x, x, B=B, H=H, W=W, Cin=cin, Cout=cout, ksize=ksize, stride=stride, prologue=prologue,
                                          epilogue=epilogue, impl=IMPL, w_tc=None, res=res, res_pool=res_pool, keep_prob=keep_prob,
                                          dropout_seed=dropout_seed, query="layout"))
    w16, ws16, bias, w_tc = _conv_weights(scope, ksize, cin * mult, cout, x.device, skip_scope, kskip, gated, pass_major)
    OH, OW = -(-H // stride), -(-W // stride)
    cy = cout // 2 if epilogue == native.EPI_GATE_RES else cout
    if query:
        return native.conv_fwd(x, w16, bias, x, B=B, H=H, W=W, Cin=cin, Cout=cout, ksize=ksize, stride=stride,
                               prologue=prologue, epilogue=epilogue, impl=IMPL, w_tc=w_tc, skip=skip, w_skip=ws16, res=res,
                               res_pool=res_pool, keep_prob=keep_prob, dropout_seed=dropout_seed, query=True)
    y = torch.empty((B, OH, OW, cy), device=x.device,
                    dtype=torch.float32 if epilogue == native.EPI_TANH else torch.bfloat16)
    rec = None
    if _PROFILE is not None:
        nbytes = x.numel() * x.element_size() + y.numel() * y.element_size()
        nbytes += skip.numel() * 2 if skip is not None else 0
        nbytes += res.numel() * res.element_size() if res is not None else 0
        flops = 2 * B * OH * OW * cout * (ksize * ksize * cin * mult + kskip)
        ksteps = ksize * ksize * (-(-cin * mult // 16)) + kskip // 16
        rec = dict(name=scope, kind="conv", out_hw=(OH, OW), K_tap=cin * mult, N=cout, stride=stride,
                   alg_bytes=nbytes, flops=flops, mma_clk=_mma_clk(B * (-(-OH * OW // 128)), ksteps, cout) if cin > 1 and cout >= 8 else 0.0)
    _profiled(rec, lambda: native.conv_fwd(
        x, w16, bias, y, B=B, H=H, W=W, Cin=cin, Cout=cout, ksize=ksize, stride=stride, prologue=prologue,
        epilogue=epilogue, impl=IMPL, w_tc=w_tc, skip=skip, w_skip=ws16, res=res, res_pool=res_pool,
        keep_prob=keep_prob, dropout_seed=dropout_seed))
    return y


# --------------------------------------------------------------------------- public layers
def conv_layer(inputs, kernel_size, stride, num_features, idx, nonlinearity=None):
    """reference :57-68 -- tf.nn.conv2d 'SAME' + bias_add (+ nonlinearity)"""
    y = _fused_conv(inputs, '{0}_conv'.format(idx), kernel_size, stride, num_features, native.PRO_NONE,
                    native.EPI_BIAS)
    if nonlinearity is not None:
        y = nonlinearity(y)
    return y


def transpose_conv_layer(inputs, kernel_size, stride, num_features, idx, nonlinearity=None):
    """reference :70-87 -- tf.nn.conv2d_transpose 'SAME' + bias_add, output [B, sH, sW, num_features]"""
    if kernel_size != 3 or stride != 2:
        raise ValueError("transpose_conv_layer: only kernel_size=3, stride=2 (the reference's only use) is built")
    x = native.to_bf16(inputs)
    B, H, W, cin = int_shape(x)
    scope = '{0}_trans_conv'.format(idx)
    w = _variable(scope + '/weights', [kernel_size, kernel_size, num_features, cin], device=x.device)
    b = _variable(scope + '/biases', [num_features], device=x.device)

    def make():
        # TF [kh,kw,Cout,Cin] -> GEMM-B [tap, Cin, Cout]
        w16 = w.permute(0, 1, 3, 2).reshape(9, cin, num_features).to(torch.bfloat16).contiguous()
        w_tc = native.pack_conv_weights_tc(w16, None, 3, cin, 0, num_features, False) if cin % 16 == 0 else None
        return w16, b.contiguous(), w_tc

    w16, bias, w_tc = VARIABLES.derived(("tconv", scope), make)
    y = torch.empty((B, 2 * H, 2 * W, num_features), device=x.device, dtype=torch.bfloat16)
    rec = None
    if _PROFILE is not None:
        rec = dict(name=scope, kind="tconv", out_hw=(2 * H, 2 * W), K_tap=cin, N=num_features, stride=2,
                   alg_bytes=(x.numel() + y.numel()) * 2, flops=2 * B * H * W * 9 * cin * num_features,
                   mma_clk=_mma_clk(B * (-(-H * W // 128)), 9 * (-(-cin // 16)), num_features))
    _profiled(rec, lambda: native.tconv_fwd(x, w16, bias, y, B=B, H=H, W=W, Cin=cin, Cout=num_features, impl=IMPL,
                                            w_tc=w_tc))
    if TAPE is not None:
        TAPE.append(dict(kind="tconv", scope=scope, x=x, out=y))
    if nonlinearity is not None:
        y = nonlinearity(y)
    return y


def fc_layer(inputs, hiddens, idx, nonlinearity=None, flat=False):
    """reference :89-104 -- matmul + bias on [rows, dim].  (The reference's flat=True and
    nonlinearity branches are dead/buggy, :92-94,:102-103; flat is honoured, nonlinearity applied.)"""
    x = native.to_bf16(inputs)
    if flat:
        x = x.reshape(x.shape[0], -1)
    rows, dim = x.shape
    scope = '{0}_fc'.format(idx)
    # run it as a 1x1 convolution over a [1, rows, 1, dim] image; the variable keeps TF's [dim, hiddens] shape
    w = _variable(scope + '/weights', [dim, hiddens], device=x.device)
    b = _variable(scope + '/biases', [hiddens], device=x.device)

    def make():
        return w.reshape(1, dim, hiddens).to(torch.bfloat16).contiguous(), b.contiguous()

    w16, bias = VARIABLES.derived(("fc", scope), make)
    y = torch.empty((1, rows, 1, hiddens), device=x.device, dtype=torch.bfloat16)
    native.conv_fwd(x.reshape(1, rows, 1, dim), w16, bias, y, B=1, H=rows, W=1, Cin=dim, Cout=hiddens, ksize=1,
                    stride=1, prologue=native.PRO_NONE, epilogue=native.EPI_BIAS, impl=native.IMPL_SIMT)
    y = y.reshape(rows, hiddens)
    if nonlinearity is not None:
        y = nonlinearity(y)
    return y


def nin(x, num_units, idx):
    """ a network in network layer (1x1 CONV) -- reference :106-111 """
    s = int_shape(x)
    x = x.reshape([int(np.prod(s[:-1])), s[-1]])
    x = fc_layer(x, num_units, idx)
    return x.reshape(s[:-1] + [num_units])


def res_block(x, a=None, filter_size=16, nonlinearity=concat_elu, keep_p=1.0, stride=1, gated=False,
              name="resnet", dropout_seed=0):
    """reference :129-165.  Two fused launches (see module docstring)."""
    pro = _PROLOGUE_OF.get(nonlinearity)
    if pro is None:
        return _res_block_unfused(x, a, filter_size, nonlinearity, keep_p, stride, gated, name)
    # inference on the boundary image as it arrives (uint8 [B,H,W,1]): the head conv and the gated conv_2's one-channel shortcut
    # read it directly (fn_conv_desc.x_u8 / res_u8) -- no cast launch, no bf16 copy of the input
    direct_u8 = (x.dtype == torch.uint8 and x.shape[3] == 1 and TAPE is None and IMPL == native.IMPL_AUTO and gated and stride == 1
                 and a is None and filter_size == 8 and pro == native.PRO_CONCAT_ELU and keep_p >= 1.0 and x.shape[1] >= 16 and x.shape[2] >= 8
                 and native.level5_supported(16, 32, 64) and os.environ.get("FLOWNET_WG", "1") != "0")
    orig_x = x if direct_u8 else native.to_bf16(x)
    cin = orig_x.shape[3]
    if a is not None and cin == 1:
        raise ValueError("res_block: a skip input on a 1-channel block is not supported")
    x_1 = _fused_conv(orig_x, name + '_conv_1_conv', 3, stride, filter_size,
                      native.PRO_NONE if cin == 1 else pro, native.EPI_BIAS,
                      skip=a, skip_scope=(name + '_nin_fc') if a is not None else None)
    pooled = orig_x.shape[2] > x_1.shape[2]
    out = _fused_conv(x_1, name + '_conv_2_conv', 3, 1, filter_size * (2 if gated else 1), pro,
                      native.EPI_GATE_RES if gated else native.EPI_RES, res=orig_x, res_pool=pooled,
                      keep_prob=keep_p, dropout_seed=dropout_seed)
    if TAPE is not None:
        TAPE.append(dict(kind="block", name=name, x=orig_x, a=native.to_bf16(a) if a is not None else None, x_1=x_1,
                         out=out, F=filter_size, stride=stride, gated=gated, pro=pro, keep_p=keep_p, seed=dropout_seed,
                         pooled=pooled))
    return out


def _res_block_unfused(x, a, filter_size, nonlinearity, keep_p, stride, gated, name):
    raise ValueError("res_block: only the reference's four nonlinearities (set_nonlinearity) are supported")


# --------------------------------------------------------------------------- the 8x16 level in one launch
LEVEL5_FUSED = os.environ.get("FLOWNET_L5", "1") != "0"


def _level5_fusable(x, nr_res_blocks, keep_prob, nonlinearity, gated, filter_size):
    """fn_level5_fwd takes the reference's flags at the 128x256 grid (x = resnet_4_0's output [B,16,32,64]) in inference."""
    return (LEVEL5_FUSED and TAPE is None and IMPL == native.IMPL_AUTO and nr_res_blocks == 1 and keep_prob >= 1.0 and
            nonlinearity is concat_elu and gated and filter_size == 128 and
            native.level5_supported(x.shape[1], x.shape[2], x.shape[3]))


# two CTAs per image (thread-block cluster, csrc/conv_l5c.cu): FLOWNET_L5_SPLIT=0 goes back to one CTA per image
LEVEL5_SPLIT = os.environ.get("FLOWNET_L5_SPLIT", "1") != "0"


def _level5_weights(dev, split):
    """The five tensor-core weight images and biases of the 8x16 level, concatenated once per variable version (split: the
    per-rank form of the two-CTA kernel); variables are created in the reference's order."""
    names = ("resnet_5_downsample_conv_1_conv", "resnet_5_downsample_conv_2_conv", "resnet_5_0_conv_1_conv",
             "resnet_5_0_conv_2_conv")
    shapes = ((128, 128, False), (256, 256, True), (256, 128, False), (256, 256, True))
    parts = [_conv_weights(n, 3, k, c, dev, gated=g) for n, (k, c, g) in zip(names, shapes)]
    scope = "up_conv_1_trans_conv"
    wt = _variable(scope + '/weights', [3, 3, 64, 128], device=dev)
    bt = _variable(scope + '/biases', [64], device=dev)

    def make_t():
        w16 = wt.permute(0, 1, 3, 2).reshape(9, 128, 64).to(torch.bfloat16).contiguous()
        return w16, bt.contiguous(), native.pack_conv_weights_tc(w16, None, 3, 128, 0, 64, False)

    tpart = VARIABLES.derived(("tconv", scope), make_t)

    def make():
        wstream = torch.cat([pt[3] for pt in parts] + [tpart[2]])
        bias = torch.cat([pt[2] for pt in parts] + [tpart[1]]).to(torch.float32).contiguous()
        return wstream, bias

    def make_split():
        # rank r computes output columns [r F/2, (r+1) F/2) of every convolution (gated: those u columns, then the matching v's)
        ws, bs = [], []
        for r in range(2):
            for (w16, _, b, _), (k, c, g) in zip(parts, shapes):
                f = c // 2 if g else c
                cols = torch.arange(r * f // 2, (r + 1) * f // 2, device=dev)
                if g:
                    cols = torch.cat([cols, cols + f])
                ws.append(native.pack_conv_weights_tc(w16[:, :, cols].contiguous(), None, 3, k, 0, cols.numel(), g))
                bs.append(b[cols])
            ws.append(native.pack_conv_weights_tc(tpart[0][:, :, r * 32:(r + 1) * 32].contiguous(), None, 3, 128, 0, 32, False))
            bs.append(tpart[1][r * 32:(r + 1) * 32])
        return torch.cat(ws), torch.cat(bs).to(torch.float32).contiguous()

    return VARIABLES.derived(("level5", split), make_split if split else make)


def _level5_fused(x):
    """resnet_5_downsample, resnet_5_0 and up_conv_1 (reference :207-215) in one launch."""
    dev = x.device
    # two CTAs per image while the clusters fit the GPU in one wave (measured: 52 vs 67 us at batch 64, 204 vs 135 us at 256)
    split = LEVEL5_SPLIT and 2 * x.shape[0] <= torch.cuda.get_device_properties(dev).multi_processor_count
    wstream, bias = _level5_weights(dev, split)
    x = native.to_bf16(x)
    B = x.shape[0]
    y = torch.empty((B, 16, 32, 64), device=dev, dtype=torch.bfloat16)
    rec = None
    if _PROFILE is not None:
        # algorithmic bytes: x read by the stride-2 conv and by the pooled shortcut counts once, y once; flops: the five convs
        macs = B * 128 * (9 * 128 * 128 + 9 * 256 * 256 + 9 * 256 * 128 + 9 * 256 * 256 + 9 * 128 * 64)
        # per image: one tile; with the channel split every CTA runs the same k-steps at half the width
        w = (64, 128, 64, 128, 32) if split else (128, 256, 128, 256, 64)
        k = (72, 144, 144, 144, 72)
        clk = sum(_mma_clk(B * (2 if split else 1), ks, n) for ks, n in zip(k, w))
        rec = dict(name="level5_fused(resnet_5_downsample+resnet_5_0+up_conv_1)", kind="fused", out_hw=(16, 32), K_tap=256, N=256,
                   stride=1, alg_bytes=(x.numel() + y.numel()) * 2, flops=2 * macs, mma_clk=clk)
    _profiled(rec, lambda: native.level5_fwd(x, wstream, bias, y, split=split))
    return y


# --------------------------------------------------------------------------- the 16x32 level in two launches
LEVEL4_FUSED = os.environ.get("FLOWNET_L4", "1") != "0"
# clusters of two CTAs per image: used while they fit the GPU in at most this many waves
LEVEL4_MAX_WAVES = float(os.environ.get("FLOWNET_L4_WAVES", "1"))


def _level4_fusable(x, nr_res_blocks, keep_prob, nonlinearity, gated, filter_size):
    """fn_level4_fwd takes the reference's flags at the 128x256 grid (x = resnet_3_0's output [B,32,64,32]) in inference;
    it runs together with the fused 8x16 level."""
    if not (LEVEL4_FUSED and LEVEL5_FUSED and TAPE is None and IMPL == native.IMPL_AUTO and nr_res_blocks == 1 and keep_prob >= 1.0
            and nonlinearity is concat_elu and gated and filter_size == 64 and tuple(x.shape[1:]) == (32, 64, 32)
            and native.level5_supported(16, 32, 64)):
        return False
    sms = torch.cuda.get_device_properties(x.device).multi_processor_count
    return 2 * x.shape[0] <= LEVEL4_MAX_WAVES * sms


def _tconv_parts(scope, cin, cout, dev):
    wt = _variable(scope + '/weights', [3, 3, cout, cin], device=dev)
    bt = _variable(scope + '/biases', [cout], device=dev)

    def make_t():
        w16 = wt.permute(0, 1, 3, 2).reshape(9, cin, cout).to(torch.bfloat16).contiguous()
        return w16, bt.contiguous(), native.pack_conv_weights_tc(w16, None, 3, cin, 0, cout, False)

    return VARIABLES.derived(("tconv", scope), make_t)


def _level4_down_weights(dev):
    names = ("resnet_4_downsample_conv_1_conv", "resnet_4_downsample_conv_2_conv", "resnet_4_0_conv_1_conv", "resnet_4_0_conv_2_conv")
    shapes = ((64, 64, False), (128, 128, True), (128, 64, False), (128, 128, True))
    parts = [_conv_weights(n, 3, k, c, dev, gated=g) for n, (k, c, g) in zip(names, shapes)]

    def make():
        return torch.cat([pt[3] for pt in parts]), torch.cat([pt[2] for pt in parts]).to(torch.float32).contiguous()

    return VARIABLES.derived(("level4", "down"), make)


def _level4_down(x):
    """resnet_4_downsample and resnet_4_0 (reference :202-205) in one launch; variables in the reference's order."""
    dev = x.device
    wstream, bias = _level4_down_weights(dev)
    x = native.to_bf16(x)
    B = x.shape[0]
    y = torch.empty((B, 16, 32, 64), device=dev, dtype=torch.bfloat16)
    rec = None
    if _PROFILE is not None:
        macs = B * 512 * 9 * (64 * 64 + 128 * 128 + 128 * 64 + 128 * 128)
        rec = dict(name="level4_down(resnet_4_downsample+resnet_4_0)", kind="fused", out_hw=(16, 32), K_tap=128, N=128, stride=1,
                   alg_bytes=(x.numel() + y.numel()) * 2, flops=2 * macs,
                   mma_clk=sum(_mma_clk(4 * B, ks, n) for ks, n in ((36, 64), (72, 128), (72, 64), (72, 128))))
    _profiled(rec, lambda: native.level4_fwd(x, wstream, bias, y))
    return y


def _level4_up_weights(dev):
    c1 = _conv_weights("resnet_up_1_0_conv_1_conv", 3, 128, 64, dev, skip_scope="resnet_up_1_0_nin_fc", kskip_eff=128)
    c2 = _conv_weights("resnet_up_1_0_conv_2_conv", 3, 128, 128, dev, gated=True)
    tp = _tconv_parts("up_conv_2_trans_conv", 64, 32, dev)

    def make():
        return torch.cat([c1[3], c2[3], tp[2]]), torch.cat([c1[2], c2[2], tp[1]]).to(torch.float32).contiguous()

    return VARIABLES.derived(("level4", "up"), make)


def _level4_up(z, a):
    """resnet_up_1_0 (with the nin on the skip a = resnet_4_0's output) and up_conv_2 (reference :217-222) in one launch.
    Variable creation order of the reference: up_conv_1 (created by the level-5 kernel's wrapper), the block's conv_1, nin,
    conv_2, then up_conv_2."""
    dev = z.device
    wstream, bias = _level4_up_weights(dev)
    B = z.shape[0]
    y = torch.empty((B, 32, 64, 32), device=dev, dtype=torch.bfloat16)
    rec = None
    if _PROFILE is not None:
        macs = B * 512 * (9 * 128 * 64 + 128 * 64 + 9 * 128 * 128 + 9 * 64 * 32)
        rec = dict(name="level4_up(resnet_up_1_0+up_conv_2)", kind="fused", out_hw=(32, 64), K_tap=128, N=128, stride=1,
                   alg_bytes=(z.numel() + a.numel() + y.numel()) * 2, flops=2 * macs,
                   mma_clk=sum(_mma_clk(4 * B, ks, n) for ks, n in ((80, 64), (72, 128), (36, 32))))
    _profiled(rec, lambda: native.level4_fwd(z, wstream, bias, y, skip=a))
    return y


# levels 4 and 5 in ONE launch (csrc/conv_l4.cu MODE_ALL): the stages of _level4_down, _level5_fused (two-CTA form) and
# _level4_up inside one cluster kernel, bit-identical to the three launches.  Opt-in (FLOWNET_L45=1): measured at batch 64 it
# is no faster back to back (106.0 vs 106.6 us, tools/ab_level45.py) and 15 us SLOWER inside the forward (0.5363 vs 0.5216 ms
# per step) -- inside a CUDA graph the two launches it removes cost far less than the hand-overs it adds (the next half's
# weight ring starts cold behind a cluster barrier, its operands are re-read through L2)
LEVEL45_MERGED = os.environ.get("FLOWNET_L45", "0") == "1"


def _level45_fused(x3):
    """resnet_4_downsample ... up_conv_2 (reference :202-222), twelve convolutions, one launch.  Variables are created in the
    reference's order (level 4 down, level 5, level 4 up)."""
    dev = x3.device
    wd, bd = _level4_down_weights(dev)
    w5, b5 = _level5_weights(dev, True)
    wu, bu = _level4_up_weights(dev)
    x3 = native.to_bf16(x3)
    B = x3.shape[0]
    x4 = torch.empty((B, 16, 32, 64), device=dev, dtype=torch.bfloat16)
    z = torch.empty((B, 16, 32, 64), device=dev, dtype=torch.bfloat16)
    y = torch.empty((B, 32, 64, 32), device=dev, dtype=torch.bfloat16)
    rec = None
    if _PROFILE is not None:
        macs = B * 512 * 9 * (64 * 64 + 128 * 128 + 128 * 64 + 128 * 128) + \
            B * 128 * (9 * 128 * 128 + 9 * 256 * 256 + 9 * 256 * 128 + 9 * 256 * 256 + 9 * 128 * 64) + \
            B * 512 * (9 * 128 * 64 + 128 * 64 + 9 * 128 * 128 + 9 * 64 * 32)
        clk = sum(_mma_clk(4 * B, ks, n) for ks, n in ((36, 64), (72, 128), (72, 64), (72, 128), (80, 64), (72, 128), (36, 32))) + \
            sum(_mma_clk(2 * B, ks, n) for ks, n in zip((72, 144, 144, 144, 72), (64, 128, 64, 128, 32)))
        rec = dict(name="levels45_fused(resnet_4_downsample..up_conv_2)", kind="fused", out_hw=(32, 64), K_tap=256, N=256, stride=1,
                   alg_bytes=(x3.numel() + y.numel()) * 2, flops=2 * macs, mma_clk=clk)
    _profiled(rec, lambda: native.level45_fwd(x3, x4, z, wd, bd, w5, b5, wu, bu, y))
    return y, x4


# --------------------------------------------------------------------------- the 32x64 level's stride-1 convs in two launches
# measured at batch 64 (tools/ab_level3.py): 64.7 -> 60.1 us and 44.6 -> 43.4 us against the layer-by-layer launches, the
# forward 121.7 k -> 122.8 k img/s; at batch 8 the fused launches are twice as slow (eight serial tiles per CTA at 48 clk per
# N = 64 MMA are 15 k clocks per stage, which the pipelined per-layer kernels spread over 148 SMs) -- hence the batch gate
LEVEL3_FUSED = os.environ.get("FLOWNET_L3", "1") != "0"
LEVEL3_MIN_BATCH = int(os.environ.get("FLOWNET_L3_MIN_BATCH", "48"))


def _level3_fusable(x, nr_res_blocks, keep_prob, nonlinearity, gated, filter_size):
    """fn_level3_fwd takes the reference's flags at the 128x256 grid (x = resnet_2_0's output [B,64,128,16]) in inference."""
    if not (LEVEL3_FUSED and TAPE is None and IMPL == native.IMPL_AUTO and nr_res_blocks == 1 and keep_prob >= 1.0
            and nonlinearity is concat_elu and gated and filter_size == 32 and tuple(x.shape[1:]) == (64, 128, 16)
            and native.level5_supported(16, 32, 64)):
        return False
    sms = torch.cuda.get_device_properties(x.device).multi_processor_count
    return x.shape[0] >= LEVEL3_MIN_BATCH and 2 * x.shape[0] <= LEVEL4_MAX_WAVES * sms


def _level3_down(x2):
    """resnet_3_downsample and resnet_3_0 (reference :198-201): the stride-2 conv_1 as its own launch, then conv_2 and the
    whole second block in one; variables in the reference's order."""
    dev = x2.device
    x2 = native.to_bf16(x2)
    x1 = _fused_conv(x2, "resnet_3_downsample_conv_1_conv", 3, 2, 32, native.PRO_CONCAT_ELU, native.EPI_BIAS)
    names = ("resnet_3_downsample_conv_2_conv", "resnet_3_0_conv_1_conv", "resnet_3_0_conv_2_conv")
    shapes = ((64, 64, True), (64, 32, False), (64, 64, True))
    parts = [_conv_weights(n, 3, k, c, dev, gated=g) for n, (k, c, g) in zip(names, shapes)]

    def make():
        return torch.cat([pt[3] for pt in parts]), torch.cat([pt[2] for pt in parts]).to(torch.float32).contiguous()

    wstream, bias = VARIABLES.derived(("level3", "down"), make)
    B = x2.shape[0]
    y = torch.empty((B, 32, 64, 32), device=dev, dtype=torch.bfloat16)
    rec = None
    if _PROFILE is not None:
        macs = B * 2048 * 9 * (64 * 64 + 64 * 32 + 64 * 64)
        rec = dict(name="level3_down(resnet_3_downsample.conv_2+resnet_3_0)", kind="fused", out_hw=(32, 64), K_tap=64, N=64, stride=1,
                   alg_bytes=(x1.numel() + x2.numel() + y.numel()) * 2, flops=2 * macs,
                   mma_clk=sum(_mma_clk(16 * B, 36, n) for n in (64, 32, 64)))
    _profiled(rec, lambda: native.level3_fwd(x1, x2, wstream, bias, y, up=False))
    return y


def _level3_up(z, a):
    """resnet_up_2_0 and up_conv_3 (reference :223-228): conv_1 with the nin on the skip a = resnet_3_0's output as its own
    launch, then conv_2 and the transposed conv in one."""
    dev = z.device
    x1 = _fused_conv(z, "resnet_up_2_0_conv_1_conv", 3, 1, 32, native.PRO_CONCAT_ELU, native.EPI_BIAS, skip=a,
                     skip_scope="resnet_up_2_0_nin_fc")
    c2 = _conv_weights("resnet_up_2_0_conv_2_conv", 3, 64, 64, dev, gated=True)
    tp = _tconv_parts("up_conv_3_trans_conv", 32, 16, dev)

    def make():
        return torch.cat([c2[3], tp[2]]), torch.cat([c2[2], tp[1]]).to(torch.float32).contiguous()

    wstream, bias = VARIABLES.derived(("level3", "up"), make)
    B = z.shape[0]
    y = torch.empty((B, 64, 128, 16), device=dev, dtype=torch.bfloat16)
    rec = None
    if _PROFILE is not None:
        macs = B * 2048 * (9 * 64 * 64 + 9 * 32 * 16)
        rec = dict(name="level3_up(resnet_up_2_0.conv_2+up_conv_3)", kind="fused", out_hw=(64, 128), K_tap=64, N=64, stride=1,
                   alg_bytes=(x1.numel() + z.numel() + y.numel()) * 2, flops=2 * macs,
                   mma_clk=_mma_clk(16 * B, 36, 64) + _mma_clk(16 * B, 18, 16))
    _profiled(rec, lambda: native.level3_fwd(x1, z, wstream, bias, y, up=True))
    return y


# --------------------------------------------------------------------------- inference on MMA-ready activations
# "planes" form of an activation t [B,H,W,C] (include/flownet.h, csrc/conv_pl.cu): bf16 [B, 2C/8, H, W, 8] = concat_elu(t)
# in 8-channel groups.  In inference every producer's epilogue writes what its consumers read: the planes for a stride-1
# 3x3 conv (whose tensor copies then land the tensor-core A operand directly: no transform stage), the raw NHWC tensor
# for shortcuts, resampling convs and the tail.  x_1 = conv_1's output exists in planes form only.
# Opt-in (FLOWNET_PLANES=1): measured at batch 64 the planes kernels are 10-30 % faster per launch than the raw-input
# kernels, but every planes tensor is twice the bytes of its raw form and its producers' epilogues store three 16-byte
# vectors per pixel group instead of one -- the whole forward came out 6 % slower.
PLANES = os.environ.get("FLOWNET_PLANES", "0") == "1"
_PL_CHANNELS = (8, 16, 32)       # levels whose stride-1 convs have a planes kernel


def _planes_usable(x, keep_prob, nonlinearity, gated, filter_size):
    B, H, W, cin = x.shape
    return (PLANES and TAPE is None and IMPL == native.IMPL_AUTO and keep_prob >= 1.0 and nonlinearity is concat_elu and gated
            and cin == 1 and filter_size in (8, 16) and H % 16 == 0 and W % 16 == 0 and H >= 128 and W >= 128)


def _planes_empty(B, H, W, C, device):
    return torch.empty((B, 2 * C // 8, H, W, 8), device=device, dtype=torch.bfloat16)


def _conv_planes(scope, cout, epilogue, x=None, xp=None, stride=1, skip_p=None, skip_scope=None, res=None,
                 res_pool=False, want_raw=True, want_planes=False):
    """One fused 3x3 conv launch of the planes path.  Input: raw `x` [B,H,W,C] (1-channel head, stride-2 convs) or planes
    `xp`; optional planes skip; outputs (raw or None, planes or None)."""
    if xp is not None:
        B, cin, H, W = xp.shape[0], xp.shape[1] * 4, xp.shape[2], xp.shape[3]
        src, pro = xp, native.PRO_CONCAT_ELU
    else:
        B, H, W, cin = int_shape(x)
        src, pro = x, (native.PRO_NONE if cin == 1 else native.PRO_CONCAT_ELU)
    mult = _pro_mult(pro)
    kskip = skip_p.shape[1] * 8 if skip_p is not None else 0
    gated = epilogue == native.EPI_GATE_RES
    w16, ws16, bias, w_tc = _conv_weights(scope, 3, cin * mult, cout, src.device, skip_scope, kskip, gated)
    OH, OW = H // stride, W // stride
    cy = cout // 2 if gated else cout
    y = torch.empty((B, OH, OW, cy), device=src.device, dtype=torch.bfloat16) if want_raw else None
    yp = _planes_empty(B, OH, OW, cy, src.device) if want_planes else None
    rec = None
    if _PROFILE is not None:
        # algorithmic bytes are counted on the RAW form of every tensor (2 bytes per channel), so the figures compare
        # with the layer-fused byte count of SURVEY App. C; `moved_bytes` is what this launch really reads and writes
        raw_in = B * H * W * cin * 2
        nbytes = raw_in + B * OH * OW * cy * 2
        moved = (raw_in * (2 if xp is not None else 1) + (y.numel() * 2 if y is not None else 0) +
                 (yp.numel() * 2 if yp is not None else 0))
        if skip_p is not None:
            nbytes += skip_p.numel()
            moved += skip_p.numel() * 2
        if res is not None:
            nbytes += res.numel() * 2
            moved += res.numel() * 2
        flops = 2 * B * OH * OW * cout * (9 * cin * mult + kskip)
        rec = dict(name=scope, kind="conv", out_hw=(OH, OW), K_tap=cin * mult, N=cout, stride=stride,
                   alg_bytes=nbytes, moved_bytes=moved, flops=flops)
    _profiled(rec, lambda: native.conv_fwd(
        src, w16, bias, y, B=B, H=H, W=W, Cin=cin, Cout=cout, ksize=3, stride=stride, prologue=pro, epilogue=epilogue,
        impl=IMPL, w_tc=w_tc, skip=skip_p, w_skip=ws16, res=res, res_pool=res_pool, x_planes=xp is not None,
        skip_planes=skip_p is not None, y_planes=yp))
    return y, yp


def _tconv_planes(x, num_features, idx, want_planes):
    """transpose_conv_layer (reference :70-87) writing the raw output and, if asked, its planes form in the same launch"""
    B, H, W, cin = int_shape(x)
    scope = '{0}_trans_conv'.format(idx)
    w = _variable(scope + '/weights', [3, 3, num_features, cin], device=x.device)
    b = _variable(scope + '/biases', [num_features], device=x.device)

    def make():
        w16 = w.permute(0, 1, 3, 2).reshape(9, cin, num_features).to(torch.bfloat16).contiguous()
        w_tc = native.pack_conv_weights_tc(w16, None, 3, cin, 0, num_features, False) if cin % 16 == 0 else None
        return w16, b.contiguous(), w_tc

    w16, bias, w_tc = VARIABLES.derived(("tconv", scope), make)
    y = torch.empty((B, 2 * H, 2 * W, num_features), device=x.device, dtype=torch.bfloat16)
    yp = _planes_empty(B, 2 * H, 2 * W, num_features, x.device) if want_planes else None
    rec = None
    if _PROFILE is not None:
        rec = dict(name=scope, kind="tconv", out_hw=(2 * H, 2 * W), K_tap=cin, N=num_features, stride=2,
                   alg_bytes=(x.numel() + y.numel()) * 2,
                   moved_bytes=(x.numel() + y.numel() + (yp.numel() if yp is not None else 0)) * 2,
                   flops=2 * B * H * W * 9 * cin * num_features)
    _profiled(rec, lambda: native.tconv_fwd(x, w16, bias, y, B=B, H=H, W=W, Cin=cin, Cout=num_features, impl=IMPL,
                                            w_tc=w_tc, y_planes=yp))
    return y, yp


def _conv_res_planes(inputs, nr_res_blocks, filter_size):
    """conv_res (reference :167-248) for inference with the reference's flags (gated, concat_elu, keep_prob 1): the same
    launches in the same order with the same variables, the activations of the 8/16/32-channel levels in planes form.
    Deeper levels run the kernels of the general path on raw tensors."""
    conc = concat_elu

    def cap(C):
        return C in _PL_CHANNELS

    def block(x, xp, F, name, stride=1, a_p=None, a=None, out_planes=False):
        """res_block (:129-165): x raw (always needed: shortcut), xp = its planes or None; returns (raw, planes or None)"""
        cin = x.shape[3]
        if not cap(F):                 # no planes kernel at this width: the general two-launch block on raw tensors
            return res_block(x, a=a, filter_size=F, nonlinearity=conc, stride=stride, gated=True, name=name), None
        skip_scope = (name + '_nin_fc') if a_p is not None else None
        if stride == 2 or cin == 1:
            _, x1p = _conv_planes(name + '_conv_1_conv', F, native.EPI_BIAS, x=x, stride=stride, want_raw=False, want_planes=True)
        else:
            _, x1p = _conv_planes(name + '_conv_1_conv', F, native.EPI_BIAS, xp=xp, skip_p=a_p, skip_scope=skip_scope,
                                  want_raw=False, want_planes=True)
        return _conv_planes(name + '_conv_2_conv', 2 * F, native.EPI_GATE_RES, xp=x1p, res=x, res_pool=(stride == 2),
                            want_raw=True, want_planes=out_planes)

    a = []
    F = filter_size
    x, xp = native.to_bf16(inputs), None
    # res_1
    for i in range(nr_res_blocks):
        x, xp = block(x, xp, F, "resnet_1_" + str(i), out_planes=cap(F))
    # res_2 .. res_5
    fused5 = False
    for lvl in (2, 3, 4, 5):
        a.append((x, xp))
        F = 2 * F
        if lvl == 5 and _level5_fusable(x, nr_res_blocks, 1.0, conc, True, F):
            x, xp = _level5_fused(x), None
            fused5 = True
            continue
        x, xp = block(x, xp, F, "resnet_%d_downsample" % lvl, stride=2, out_planes=cap(F))
        for i in range(nr_res_blocks):
            x, xp = block(x, xp, F, "resnet_%d_%d" % (lvl, i), out_planes=cap(F))
    # res_up_1 .. res_up_4
    for k in (1, 2, 3, 4):
        F = F // 2
        if not (k == 1 and fused5):
            if cap(F):
                x, xp = _tconv_planes(x, F, "up_conv_%d" % k, want_planes=True)
            else:
                x, xp = transpose_conv_layer(x, 3, 2, F, "up_conv_%d" % k), None
        for i in range(nr_res_blocks):
            ar, ap = a[-k] if i == 0 else (None, None)
            x, xp = block(x, xp, F, "resnet_up_%d_%d" % (k, i), a_p=ap, a=ar,
                          out_planes=cap(F) and i + 1 < nr_res_blocks)
    return _fused_conv(x, 'last_conv_conv', 3, 1, 2, native.PRO_NONE, native.EPI_TANH)


def conv_res(inputs, nr_res_blocks=1, keep_prob=1.0, nonlinearity_name='concat_elu', gated=True, filter_size=8,
             dropout_seed=0):
    """Builds conv part of net -- reference :167-248.
    Args:
      inputs: boundary images [B,H,W,1] (uint8 / float32 / bfloat16, CUDA)
      keep_prob: dropout keep probability (1.0 = inference)
      filter_size: base width; hard-coded to 8 in the reference (:174), exposed for wider variants
    Returns: [B,H,W,2] float32 velocity field in (-1, 1)
    """
    nonlinearity = set_nonlinearity(nonlinearity_name)
    if _planes_usable(inputs, keep_prob, nonlinearity, gated, filter_size):
        return _conv_res_planes(inputs, nr_res_blocks, filter_size)
    seed = [int(dropout_seed)]

    def block(x, a=None, stride=1, name=""):
        seed[0] += 1
        return res_block(x, a=a, filter_size=filter_size, nonlinearity=nonlinearity, keep_p=keep_prob, stride=stride,
                         gated=gated, name=name, dropout_seed=seed[0])

    # store for as
    a = []
    # res_1
    x = inputs if inputs.dtype == torch.uint8 else native.to_bf16(inputs)      # uint8: see res_block
    for i in range(nr_res_blocks):
        x = block(x, name="resnet_1_" + str(i))
    # res_2 .. res_5
    fused5 = fused4 = fused3 = up3_done = fused45 = False
    for lvl in (2, 3, 4, 5):
        a.append(x)
        filter_size = 2 * filter_size
        if lvl == 3 and _level3_fusable(x, nr_res_blocks, keep_prob, nonlinearity, gated, filter_size):
            # resnet_3_downsample + resnet_3_0 in two launches (csrc/conv_l3.cu)
            x = _level3_down(x)
            seed[0] += 2
            fused3 = True
            continue
        if lvl == 4 and _level4_fusable(x, nr_res_blocks, keep_prob, nonlinearity, gated, filter_size):
            if LEVEL45_MERGED and LEVEL5_SPLIT:
                # levels 4 and 5 entirely (down half, 8x16 level, up half incl. up_conv_2) in one launch
                x, x4 = _level45_fused(x)
                a.append(x4)
                seed[0] += 5
                fused45 = True
                break
            # resnet_4_downsample + resnet_4_0 in one launch (csrc/conv_l4.cu); the matching up half runs after level 5
            x = _level4_down(x)
            seed[0] += 2
            fused4 = True
            continue
        if lvl == 5 and _level5_fusable(x, nr_res_blocks, keep_prob, nonlinearity, gated, filter_size):
            # the whole 8x16 level (down block, block, up_conv_1) in one launch: csrc/conv_l5.cu
            x = _level5_fused(x)
            seed[0] += 2                           # the two blocks it contains keep their dropout-seed numbers
            fused5 = True
            continue
        x = block(x, stride=2, name="resnet_%d_downsample" % lvl)
        for i in range(nr_res_blocks):
            x = block(x, name="resnet_%d_%d" % (lvl, i))
    # res_up_1 .. res_up_4
    if fused45:
        filter_size = 2 * filter_size          # the loop left before level 5 doubled it
    for k in (1, 2, 3, 4):
        filter_size = int(filter_size / 2)
        if k == 1 and fused45:
            continue                            # resnet_up_1_0 and up_conv_2 ran inside the merged launch
        if k == 1 and fused4 and fused5:
            # resnet_up_1_0 + up_conv_2 in one launch; level 3's loop turn then skips its transposed conv
            x = _level4_up(x, a[-1])
            seed[0] += 1
            continue
        if not ((k == 1 and fused5) or (k == 2 and ((fused4 and fused5) or fused45)) or (k == 3 and up3_done)):
            x = transpose_conv_layer(x, 3, 2, filter_size, "up_conv_%d" % k)
        if k == 2 and fused3 and nr_res_blocks == 1:
            # resnet_up_2_0 + up_conv_3 in two launches; level 2's loop turn then skips its transposed conv
            x = _level3_up(x, a[-2])
            seed[0] += 1
            up3_done = True
            continue
        for i in range(nr_res_blocks):
            x = block(x, a=a[-k] if i == 0 else None, name="resnet_up_%d_%d" % (k, i))
    # last_conv + tanh (reference :242-243), one launch
    y = _fused_conv(x, 'last_conv_conv', 3, 1, 2, native.PRO_NONE, native.EPI_TANH)
    if TAPE is not None:
        TAPE.append(dict(kind="last", x=x, out=y))
    return y
