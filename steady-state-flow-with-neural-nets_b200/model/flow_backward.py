"""Backward pass + TF1 Adam for flow_net -- what `AdamOptimizer(lr).minimize(loss)` builds in the
reference (/root/reference/model/flow_net.py:44-46) for the graph of flow_architecture.py:129-248.

Every gradient is computed by kernels of libflownet.so: the data gradients of the convolutions are the FORWARD tensor-core
kernels on the gradient tensor (stride-1 conv: flipped, transposed weights; stride-2 conv <-> transposed conv with the
same weights) with the prologue's adjoint fused into their epilogues, the pre-gate tensor is recomputed instead of
stored (gate adjoint in the recomputing launch's epilogue), weight and bias gradients run on tcgen05 (csrc/wgrad_tc.cu).
Activation gradients are bf16, parameter gradients / Adam state fp32 in flat buffers.  Data-parallel training: gradients
of a SUM loss simply add across ranks; the flat gradient is reduced in two buckets -- the decoder + deepest
level as soon as their part of the backward pass is done, overlapping the encoder's backward (BackwardPass, bucket_split)
-- or in one call after the pass (allreduce_gradients).
"""
from __future__ import annotations

import torch

import flownet_native as native
import model.flow_architecture as arch


class FlatParams:
    """All variables re-homed as views of ONE flat fp32 buffer (same names / shapes / TF layouts), with
    matching flat gradient and Adam-slot buffers."""

    def __init__(self):
        vs = arch.VARIABLES
        if not vs.vars:
            raise RuntimeError("run one forward first: variables are created lazily, as in the reference")
        n = sum(v.numel() for v in vs.vars.values())
        dev = next(iter(vs.vars.values())).device
        self.param = torch.empty(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        self.offsets = {}
        off = 0
        for name, v in list(vs.vars.items()):
            k = v.numel()
            self.param[off:off + k].copy_(v.reshape(-1))
            vs.vars[name] = self.param[off:off + k].view(v.shape)
            self.offsets[name] = (off, k, tuple(v.shape))
            off += k
        self.n = n
        self.step = 0
        vs.touch()

    def grad_view(self, name):
        off, k, shape = self.offsets[name]
        return self.grad[off:off + k].view(shape)


_ZERO_BIAS = {}


def _zeros_bias(n, device):
    """the zero bias of the bias-free backward convs: one persistent tensor per (width, device), never written"""
    key = (int(n), str(device))
    z = _ZERO_BIAS.get(key)
    if z is None:
        z = _ZERO_BIAS[key] = torch.zeros(n, dtype=torch.float32, device=device)
    return z


class _BwdWeights:
    """B operand of a backward conv, made on demand from an fp32 variable viewed as [T][A][B]: logical weights
    W'[tap][k][n] = src[flip ? T-1-tap : tap][k][n] (or src[..][n][k] with transpose).  On the tensor-core path one
    native kernel writes the packed image straight from the variable (fn_pack_conv_weights_tc_f32); the bf16 [T,K,N]
    tensor the CUDA-core path reads is only built when that path runs."""

    def __init__(self, src, ksize, K, N, flip=False, transpose=False):
        self.src = src.contiguous()
        self.ksize, self.K, self.N, self.flip, self.transpose = ksize, K, N, flip, transpose

    def tc(self):
        if self.K % 16 and self.K != 8:
            return None
        return arch.weight_prep(
            lambda: native.pack_conv_weights_tc_f32(self.src, self.ksize, self.K, self.N, self.flip, self.transpose))

    def w16(self):
        T = self.ksize * self.ksize
        w = self.src.reshape(T, self.N, self.K) if self.transpose else self.src.reshape(T, self.K, self.N)
        if self.flip:
            w = torch.flip(w, dims=(0,))
        if self.transpose:
            w = w.permute(0, 2, 1)
        return w.to(torch.bfloat16).contiguous()


# the nin skip's data gradient: one CUDA-core launch (1x1 conv + prologue adjoint) on the wide, shallow levels, the 1x1 conv on
# the tensor cores + fn_prologue_bwd where the channel count makes the CUDA-core product the cost (tools/ab_nin_bwd.py, batch
# 256: F = 8 157 vs 265 us, F = 16 100 vs 121, F = 32 102 vs 80, F = 64 126 vs 38)
NIN_BWD_FUSED_WIDTHS = (8, 16)


def _pad_channels(w, axis, n):
    """the fp32 variable w zero-padded to n entries along `axis` (the 2-channel tail / 1-channel head run as 8-channel launches)"""
    shape = list(w.shape)
    shape[axis] = n
    out = torch.zeros(shape, dtype=torch.float32, device=w.device)
    out.narrow(axis, 0, w.shape[axis]).copy_(w)
    return out


def _dgrad_weights_s1(w, ksize):
    """HWIO [k,k,Keff,N] -> GEMM-B of the data-gradient conv: W'[tap'][n][keff] = W[T-1-tap'][keff][n]."""
    keff, n = w.shape[2], w.shape[3]
    return _BwdWeights(w, ksize, n, keff, flip=True, transpose=True)


def _conv_call(x, wts, y, **kw):
    """native.conv_fwd with the weights of a _BwdWeights: tensor cores if a kernel exists (the `w` argument is then a
    placeholder nobody reads), else the CUDA-core path on the bf16 [T,K,N] tensor.  Returns False if `tc_only` and no
    tensor-core kernel exists."""
    tc_only = kw.pop("tc_only", False)
    bias = _zeros_bias(wts.N, x.device)
    w_tc = wts.tc() if arch.IMPL != native.IMPL_SIMT else None
    if w_tc is not None and native.conv_fwd(x, w_tc, bias, y, w_tc=w_tc, query=True, **kw):
        native.conv_fwd(x, w_tc, bias, y, w_tc=w_tc, **kw)
        return True
    if tc_only:
        return False
    native.conv_fwd(x, wts.w16(), bias, y, w_tc=None, **kw)
    return True


def _conv_plain(x, wts, cout, ksize, stride, device, out=None):
    """bias-free conv through the forward kernels (tcgen05 when the shape allows); written into `out` when given."""
    B, H, W, cin = x.shape
    OH, OW = -(-H // stride), -(-W // stride)
    y = out if out is not None else torch.empty((B, OH, OW, cout), dtype=torch.bfloat16, device=device)
    assert tuple(y.shape) == (B, OH, OW, cout) and y.is_contiguous()
    if not isinstance(wts, _BwdWeights):            # a ready bf16 [T, K, N] tensor
        w_tc = (arch.weight_prep(lambda: native.pack_conv_weights_tc(wts, None, ksize, cin, 0, cout, False))
                if (cin % 16 == 0 or cin == 8) else None)
        native.conv_fwd(x, wts, _zeros_bias(cout, device), y, B=B, H=H, W=W, Cin=cin, Cout=cout, ksize=ksize, stride=stride,
                        prologue=native.PRO_NONE, epilogue=native.EPI_BIAS, impl=arch.IMPL, w_tc=w_tc)
        return y
    _conv_call(x, wts, y, B=B, H=H, W=W, Cin=cin, Cout=cout, ksize=ksize, stride=stride, prologue=native.PRO_NONE,
               epilogue=native.EPI_BIAS, impl=arch.IMPL)
    return y


def _conv_dgrad_adjoint(g, wts, cout, z, pro, device, out=None, accumulate=False, keep_prob=1.0, seed=0):
    """d(z) (+)= adjoint_of_prologue(conv3x3(g, W'), z): the data-gradient conv with the prologue's adjoint (and the
    forward's dropout mask) applied in its epilogue when the persistent kernel takes the shape (FN_EPI_PRO_BWD), else the
    conv followed by fn_prologue_bwd.  Returns the d(z) tensor."""
    B, H, W, cin = g.shape
    if out is None:
        out = torch.empty(tuple(z.shape), dtype=torch.bfloat16, device=device)
        accumulate = False
    if _conv_call(g, wts, out, tc_only=True, B=B, H=H, W=W, Cin=cin, Cout=cout, ksize=3, stride=1, prologue=native.PRO_NONE,
                  epilogue=native.EPI_PRO_BWD, impl=arch.IMPL, res=z, keep_prob=keep_prob, dropout_seed=seed, adj_prologue=pro,
                  adj_accumulate=accumulate):
        return out
    dA = _conv_plain(g, wts, cout, 3, 1, device)
    return native.prologue_bwd(dA, z, pro, out=out, accumulate=accumulate, keep_prob=keep_prob, seed=seed)


def _tconv_dgrad_adjoint(g, wts, cout, z, pro, device, out, accumulate):
    """d(z) (+)= adjoint_of_prologue(conv2d_transpose(g, W), z): the data gradient of a stride-2 conv with the prologue's
    adjoint in the epilogue of the sub-pixel kernel (FN_EPI_PRO_BWD) when it takes the shape, else two launches."""
    B, H, W, cin = g.shape
    w_tc = wts.tc() if (cin % 16 == 0 and arch.IMPL != native.IMPL_SIMT) else None
    if w_tc is not None:
        kw = dict(B=B, H=H, W=W, Cin=cin, Cout=cout, impl=arch.IMPL, w_tc=w_tc, epilogue=native.EPI_PRO_BWD, res=z,
                  adj_prologue=pro, adj_accumulate=accumulate)
        bias = _zeros_bias(cout, device)
        if cout == cin and pro in (native.PRO_CONCAT_ELU, native.PRO_CONCAT_RELU) and native.tconv_fwd(g, w_tc, bias, out, query=True, **kw):
            native.tconv_fwd(g, w_tc, bias, out, **kw)
            return out
    dA = _tconv_plain(g, wts, cout, device)
    return native.prologue_bwd(dA, z, pro, out=out, accumulate=accumulate)


def _tconv_plain(x, wts, cout, device):
    B, H, W, cin = x.shape
    y = torch.empty((B, 2 * H, 2 * W, cout), dtype=torch.bfloat16, device=device)
    w_tc = wts.tc() if (cin % 16 == 0 and arch.IMPL != native.IMPL_SIMT) else None
    w16 = w_tc if w_tc is not None else wts.w16()      # with a tensor-core image `w` is a placeholder nobody reads
    native.tconv_fwd(x, w16, _zeros_bias(cout, device), y, B=B, H=H, W=W, Cin=cin, Cout=cout, impl=arch.IMPL, w_tc=w_tc)
    return y


class _Grads:
    """Gradient accumulators for activation tensors with more than one consumer (block outputs that
    also feed a decoder skip): the first writer assigns, later writers accumulate in the kernel."""

    def __init__(self):
        self.g = {}

    def has(self, t):
        return id(t) in self.g

    def get(self, t):
        return self.g[id(t)][1]

    def target(self, t):
        """returns (buffer, accumulate) for a kernel that writes d(t)"""
        ent = self.g.get(id(t))
        if ent is None:
            buf = torch.empty(t.shape, dtype=torch.bfloat16, device=t.device)
            self.g[id(t)] = (t, buf)          # keep `t` alive so ids stay unique
            return buf, False
        return ent[1], True


class BackwardPass:
    """The backward pass over a recorded forward, runnable in pieces: run(records) consumes tape records in reverse order
    and may be called several times with consecutive slices of the tape (last slice first) -- flow_net.CompiledTraining
    captures the decoder / deepest-level part and the encoder part into separate CUDA graphs so that the first gradient
    bucket can be all-reduced while the second part runs.  Fills flat.grad with d(l2_loss)/d(variables) (skipped when flat
    is None); `loss` is fp32 [1]; with want_input_grad `d_input` is d loss / d boundary [B,H,W,1] bf16 -- the one data
    gradient training never needs (the first conv's), which the boundary-optimisation use of the network does (SURVEY 8f
    rank 3, README.md:31-36)."""

    def __init__(self, sflow_p, sflow, flat: FlatParams = None, want_input_grad=False):
        self.flat, self.want_input_grad = flat, want_input_grad
        self.dev = sflow_p.device
        self.grads = _Grads()
        self.loss, self.dpre = native.tanh_l2_bwd(sflow_p, sflow.to(torch.float32))
        self.d_input = None
        # parameter gradients on a stream of their own (flow_net.CompiledTraining sets it while capturing): nothing in the
        # pass reads them, so in the graph the weight / bias gradient launches leave the data-gradient chain -- at small
        # per-GPU batches, where every launch is latency, the step then costs its critical path only
        self.wstream = None
        self._wkeep = []

    def _param_grad(self, fn, *operands):
        """run fn (weight / bias gradient launches; they share native's reduction workspace, so all of them go here) on
        the parameter-gradient stream, ordered after everything issued so far"""
        W = self.wstream
        if W is None:
            fn()
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        W.wait_event(ev)
        with torch.cuda.stream(W):
            fn()
        self._wkeep.extend(operands)          # allocated on the main stream, read on W: not to be recycled before the join

    def join(self):
        """the pieces run so far are complete on the current stream once this returns (call at the end of every run slice)"""
        if self.wstream is not None:
            torch.cuda.current_stream().wait_stream(self.wstream)

    def run(self, records):
        V = arch.VARIABLES.vars
        flat, dev, grads, dpre, want_input_grad = self.flat, self.dev, self.grads, self.dpre, self.want_input_grad
        params = flat is not None

        def wgrad_into(name, x, dY, ksize, stride, pro, keep_prob=1.0, seed=0, bias=False):
            """weight gradient (and, with bias=True, the bias gradient of the same dY in the same launch)"""
            if params:
                self._param_grad(lambda: native.conv_wgrad(
                    x, dY, flat.grad_view(name + '/weights').view(ksize * ksize, -1, dY.shape[-1]), ksize=ksize, stride=stride,
                    prologue=pro, keep_prob=keep_prob, dropout_seed=seed, db=flat.grad_view(name + '/biases') if bias else None),
                    x, dY)

        def bgrad_into(name, dY):
            if params:
                self._param_grad(lambda: native.bias_grad(dY, flat.grad_view(name + '/biases')), dY)

        for rec in reversed(records):
            kind = rec["kind"]
            if kind == "last":
                x = rec["x"]                                   # [B,H,W,F] raw block output, no prologue
                cin = x.shape[3]
                # weights: [3,3,cin,2]; gradients run as 8-channel launches (dpre channels 2..7 are zero)
                if params:
                    def last_grads():
                        tmp = torch.empty((9, cin, 8), dtype=torch.float32, device=dev)
                        native.conv_wgrad(x, dpre, tmp, ksize=3, stride=1, prologue=native.PRO_NONE)
                        flat.grad_view('last_conv_conv/weights').copy_(tmp[:, :, :2].reshape(3, 3, cin, 2))
                        tmpb = torch.empty(8, dtype=torch.float32, device=dev)
                        native.bias_grad(dpre, tmpb)
                        flat.grad_view('last_conv_conv/biases').copy_(tmpb[:2])
                    self._param_grad(last_grads, x, dpre)
                w = V['last_conv_conv/weights']
                w8 = arch.weight_prep(lambda: _pad_channels(w, 3, 8))      # reads the variable only: side stream of a capture
                buf, acc = grads.target(x)
                assert not acc
                _conv_plain(dpre, _dgrad_weights_s1(w8, 3), cin, 3, 1, dev, out=buf)
            elif kind == "tconv":
                x, y, scope = rec["x"], rec["out"], rec["scope"]
                g = grads.get(y)
                cin, cout = x.shape[3], y.shape[3]
                if params:
                    self._param_grad(lambda: native.tconv_wgrad(x, g, flat.grad_view(scope + '/weights').view(9, cout, cin)), x, g)
                bgrad_into(scope, g)
                # d(x) = stride-2 'SAME' conv of d(y) with the same weights read as HWIO [k,k,Cout,Cin]
                buf, acc = grads.target(x)
                assert not acc
                _conv_plain(g, _BwdWeights(V[scope + '/weights'], 3, cout, cin), cin, 3, 2, dev, out=buf)
            elif kind == "block":
                name, x, a, x_1, out = rec["name"], rec["x"], rec["a"], rec["x_1"], rec["out"]
                F_, stride, gated, pro = rec["F"], rec["stride"], rec["gated"], rec["pro"]
                keep_p, seed, pooled = rec["keep_p"], rec["seed"], rec["pooled"]
                mult = 2 if pro in (native.PRO_CONCAT_ELU, native.PRO_CONCAT_RELU) else 1
                g = grads.get(out)
                cin = x.shape[3]
                s2 = name + '_conv_2_conv'
                n2 = 2 * F_ if gated else F_
                # ---- conv_2 (+gate): recompute the pre-gate tensor, then its adjoint
                if gated:
                    kw = dict(res=g, keep_prob=keep_p, dropout_seed=seed)
                    if arch._fused_conv(x_1, s2, 3, 1, n2, pro, native.EPI_GATE_BWD, query=True, **kw):
                        # one launch: recompute [u|v] on the tensor cores, apply the gate's adjoint in the epilogue
                        dc2 = arch._fused_conv(x_1, s2, 3, 1, n2, pro, native.EPI_GATE_BWD, **kw)
                    else:
                        c2 = arch._fused_conv(x_1, s2, 3, 1, n2, pro, native.EPI_BIAS, keep_prob=keep_p, dropout_seed=seed)
                        dc2 = native.gate_bwd(c2, g)
                else:
                    dc2 = g
                wgrad_into(s2, x_1, dc2, 3, 1, pro, keep_p, seed, bias=True)
                dx_1 = _conv_dgrad_adjoint(dc2, _dgrad_weights_s1(V[s2 + '/weights'], 3), mult * F_, x_1, pro, dev,
                                           keep_prob=keep_p, seed=seed)
                # ---- conv_1 (+nin skip)
                s1 = name + '_conv_1_conv'
                pro1 = native.PRO_NONE if cin == 1 else pro
                wgrad_into(s1, x, dx_1, 3, stride, pro1, bias=True)
                if a is not None:
                    sn = name + '_nin_fc'
                    if params:      # behind conv_1's bias gradient, i.e. on its stream
                        self._param_grad(lambda: flat.grad_view(sn + '/biases').copy_(flat.grad_view(s1 + '/biases')))
                    # the skip is zero-padded bottom/right to x_1's size: only the overlapping window contributes
                    if a.shape[1] != dx_1.shape[1] or a.shape[2] != dx_1.shape[2]:
                        raise ValueError("training with a spatially padded skip (H, W not multiples of 16) is not built")
                    if params:
                        self._param_grad(lambda: native.conv_wgrad(a, dx_1, flat.grad_view(sn + '/weights').view(1, -1, F_), ksize=1,
                                                                   stride=1, prologue=pro), a, dx_1)
                    wn = V[sn + '/weights']                                   # [mult*Ca, F]
                    buf, acc = grads.target(a)
                    if mult == 2 and F_ in NIN_BWD_FUSED_WIDTHS and wn.numel() * 4 <= 64 * 1024 and wn.is_contiguous():
                        native.nin_bwd(dx_1, wn, a, pro, out=buf, accumulate=acc)       # 1x1 conv + prologue adjoint, one launch
                    else:
                        wts = arch.weight_prep(lambda: wn.t().reshape(1, F_, -1).to(torch.bfloat16).contiguous())
                        dAa = _conv_plain(dx_1, wts, wn.shape[0], 1, 1, dev)             # 1x1 conv on the tensor cores
                        native.prologue_bwd(dAa, a, pro, out=buf, accumulate=acc)
                if cin > 1:
                    w1 = V[s1 + '/weights']                                   # [3,3,mult*cin,F]
                    buf, acc = grads.target(x)
                    if stride == 1:
                        _conv_dgrad_adjoint(dx_1, _dgrad_weights_s1(w1, 3), mult * cin, x, pro, dev, out=buf, accumulate=acc)
                    else:
                        if x.shape[1] % 2 or x.shape[2] % 2:
                            raise ValueError("training a stride-2 block on odd spatial sizes is not built")
                        _tconv_dgrad_adjoint(dx_1, _BwdWeights(w1, 3, F_, mult * cin, transpose=True), mult * cin, x, pro, dev,
                                             buf, acc)                                 # weights as [tap, F, Keff]
                    native.shortcut_bwd(g, tuple(x.shape), pooled, out=buf, accumulate=True)
                elif want_input_grad:
                    # the network input (1 channel, no prologue): conv adjoint with the output padded to 8 channels, plus
                    # the shortcut's front-padded identity (channel F-1 of g)
                    w1 = V[s1 + '/weights']                                   # [3,3,1,F]
                    w8 = arch.weight_prep(lambda: _pad_channels(w1, 2, 8))
                    dA1 = _conv_plain(dx_1, _dgrad_weights_s1(w8, 3), 8, 3, 1, dev)
                    self.d_input = d_input = dA1[..., :1].contiguous()
                    native.shortcut_bwd(g, tuple(x.shape), pooled, out=d_input, accumulate=True)


def backward(tape, sflow_p, sflow, flat: FlatParams = None, want_input_grad=False):
    """The whole backward pass in one go: returns the loss (fp32 [1]), or (loss, d loss / d boundary) with want_input_grad."""
    bp = BackwardPass(sflow_p, sflow, flat, want_input_grad)
    bp.run(tape)
    return (bp.loss, bp.d_input) if want_input_grad else bp.loss


def bucket_split(tape, flat: FlatParams, first_block="resnet_5_downsample"):
    """Where to cut the backward pass for the two-bucket gradient all-reduce: (k, off) such that run(tape[k:]) fills
    flat.grad[off:] completely -- every variable created from `first_block` on (variables are created, hence laid out,
    in forward order) -- and run(tape[:k]) fills flat.grad[:off].  With the default cut the first bucket is the deepest
    level and the decoder: ~2.2 of the 2.56 M parameters, done after ~60 % of the backward pass.  (None, None) if the
    network has no such block."""
    k = next((i for i, rec in enumerate(tape) if rec.get("kind") == "block" and rec.get("name") == first_block), None)
    key = first_block + "_conv_1_conv/weights"
    if k is None or key not in flat.offsets:
        return None, None
    return k, flat.offsets[key][0]


def allreduce_gradients(flat_grad):
    """The one collective of data-parallel training: SUM of the flat fp32 gradient over ranks.  The loss
    is a plain sum over the batch (tf.nn.l2_loss), so shard gradients ADD -- no 1/R scaling --
    and the reduced gradient equals the single-process full-batch gradient."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    return flat_grad


def adam_update(flat: FlatParams, lr, beta1=0.9, beta2=0.999, eps=1e-8, reduced=False):
    """TF1 ApplyAdam on the flat buffers, after the gradient all-reduce (reduced=True: the caller has done it)."""
    if not reduced:
        allreduce_gradients(flat.grad)
    flat.step += 1
    native.adam_step(flat.param, flat.grad, flat.m, flat.v, flat.step, lr, beta1, beta2, eps)
    arch.VARIABLES.touch()
