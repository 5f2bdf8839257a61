"""Builds the flow network -- B200-native facade with the reference's surface
(/root/reference/model/flow_net.py): flags `model`, `nr_res_blocks`, `gated_res`, `nonlinearity`
(:20-27) and functions inputs() :29, inference() :33, loss_image() :39, train() :44.

Tensors are CUDA torch tensors, NHWC: boundary [B,H,W,1] (uint8 / float32 / bfloat16, values {0,1}),
velocity field [B,H,W,2] float32.  All arithmetic runs in libflownet.so (sm_100a); there is no
CPU or PyTorch fallback.
"""
from __future__ import annotations

import contextlib
import os

import numpy as np
import torch

import flownet_native as native
import model.flow_architecture as flow_architecture
from model import flags
from model.flags import FLAGS

# Constants describing the training process (reference :20-27)
flags.DEFINE_string('model', 'res', """ model name to train """)
flags.DEFINE_integer('nr_res_blocks', 1, """ nr res blocks """)
flags.DEFINE_bool('gated_res', True, """ gated resnet or not """)
flags.DEFINE_string('nonlinearity', 'concat_elu',
                    """ nonlinearity used such as concat_elu, elu, concat_relu, relu """)
# extension: base filter count, hard-coded to 8 in the reference (flow_architecture.py:174)
flags.DEFINE_integer('filter_size', 8, """ base number of filters (reference: 8) """)


_QUEUES = {}


def inputs(batch_size, shape=(128, 256), seed=0, device="cuda", data_glob=None, shard=(0, 1)):
    """reference :29-31 returns a (boundary, sflow) batch from the TFRecord queue.  With `data_glob` (e.g.
    '../data/*.tfrecords', the reference's location) batches come from input.flow_input's shuffle queue over those
    files (shard = (rank, world): every world-th record, for data-parallel training); without a dataset (none ships, and
    there is no network) this yields the synthetic stand-in of the same shapes: boundary uint8 [B,H,W,1], sflow float32
    [B,H,W,2]."""
    if data_glob is not None:
        import input.flow_input as flow_input
        key = (data_glob, batch_size, tuple(shape), str(device))
        q = _QUEUES.get(key)
        if q is None:
            q = _QUEUES[key] = flow_input.flow_inputs(batch_size, data_glob, shape=shape, seed=seed, device=device, shard=shard)
        return q.next()
    from utils.synthetic import synth_boundary, synth_sflow
    b = synth_boundary(batch_size, shape[0], shape[1], seed=seed)
    s = synth_sflow(b)
    return torch.from_numpy(b).to(device), torch.from_numpy(s).to(device)


def _to_device(boundary):
    if isinstance(boundary, np.ndarray):
        boundary = torch.from_numpy(boundary)
    if not boundary.is_cuda:
        boundary = boundary.cuda(non_blocking=True)
    return boundary


def inference(boundary, keep_prob, dropout_seed=0):
    """reference :33-37.  boundary: [B,H,W,1]; returns sflow_p [B,H,W,2] float32 (tanh range)."""
    if FLAGS.model == "res":
        sflow_p = flow_architecture.conv_res(_to_device(boundary), nr_res_blocks=FLAGS.nr_res_blocks,
                                             keep_prob=keep_prob, nonlinearity_name=FLAGS.nonlinearity,
                                             gated=FLAGS.gated_res, filter_size=FLAGS.filter_size,
                                             dropout_seed=dropout_seed)
    else:
        # the reference falls through to an UnboundLocalError here (:34-37); say what is wrong instead
        raise ValueError("FLAGS.model=%r: only 'res' exists" % (FLAGS.model,))
    return sflow_p


def loss_image(sflow_p, sflow):
    """reference :39-42: tf.nn.l2_loss(sflow_p - sflow) = sum((p-t)^2)/2 -> float32 [1] on the device."""
    loss, _ = native.l2_loss(sflow_p, sflow, want_grad=False)
    return loss


class CompiledInference:
    """inference() for a fixed (B,H,W), captured once into a CUDA graph and replayed: the ~31 fused
    kernels of one forward are launch-bound at small batch (hard part 5).

    call(boundary) copies the batch into the static input buffer, replays, and returns the static
    float32 output buffer (valid until the next call)."""

    def __init__(self, batch_size, shape=(128, 256), keep_prob=1.0, in_dtype=torch.uint8):
        self.B, (self.H, self.W) = batch_size, shape
        self.keep_prob = keep_prob
        self.static_in = torch.zeros((batch_size, shape[0], shape[1], 1), dtype=in_dtype, device="cuda")
        self.replays = 0                               # graph replays issued so far (bench.py counts launches with it)
        self._capture()

    def _capture(self):
        """(Re-)capture the forward.  The graph has the addresses of the packed weight images baked in; those live in
        VARIABLES' derived cache, which an optimizer step / checkpoint restore clears -- so the object holds its own
        references to them and re-captures when the variables have changed since (VARIABLES.version)."""
        V = flow_architecture.VARIABLES
        # warm-up on a side stream (creates variables, packs weights, sizes the allocator pool)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                inference(self.static_in, self.keep_prob)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        n0 = native.launch_count()
        with torch.cuda.graph(self.graph):
            self.static_out = inference(self.static_in, self.keep_prob)
        self.launches_per_replay = native.launch_count() - n0
        self._version = V.version
        self._keepalive = list(V._derived.values())
        self.__dict__.pop("_pipe", None)              # the two-graph serving pipeline is rebuilt on demand

    def _fresh(self):
        if flow_architecture.VARIABLES.version != self._version:
            self._capture()

    def __call__(self, boundary):
        self._fresh()
        self.static_in.copy_(boundary, non_blocking=True)
        self.graph.replay()
        self.replays += 1
        return self.static_out

    def replay(self):
        self._fresh()
        self.graph.replay()
        self.replays += 1
        return self.static_out

    def stream_host(self, host_inputs, host_outputs):
        """Serving loop over HOST buffers with the copies overlapped with compute: step i's pinned boundary batch is
        uploaded on a copy stream while step i-1 computes and step i-2's velocity field is downloaded on a second copy
        stream.  Two captured graphs with their own static input / output buffers alternate, so uploads land directly
        in a graph's input and downloads read directly from a graph's output (no device-to-device staging copies).
        host_inputs / host_outputs: equally long sequences of pinned CPU tensors ([B,H,W,1] / float32 [B,H,W,2]).
        Returns after every output has landed on the host (the download stream is synchronised)."""
        self._fresh()
        if not hasattr(self, "_pipe"):
            # second buffer set + graph (the first is the one __init__ captured)
            in2 = torch.zeros_like(self.static_in)
            g2 = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            with torch.cuda.graph(g2):
                out2 = inference(in2, self.keep_prob)
            self._pipe = dict(
                h2d=torch.cuda.Stream(), d2h=torch.cuda.Stream(),
                din=[self.static_in, in2], dout=[self.static_out, out2], graph=[self.graph, g2],
                up=[torch.cuda.Event() for _ in range(2)], ready=[torch.cuda.Event() for _ in range(2)],
                down=[torch.cuda.Event() for _ in range(2)])
        P = self._pipe
        main = torch.cuda.current_stream()
        n = len(host_inputs)
        P["h2d"].wait_stream(main)
        for i in range(n):
            k = i & 1
            with torch.cuda.stream(P["h2d"]):
                if i >= 2:
                    P["h2d"].wait_event(P["ready"][k])         # graph k of step i-2 has consumed din[k]
                P["din"][k].copy_(host_inputs[i], non_blocking=True)
                P["up"][k].record(P["h2d"])
            main.wait_event(P["up"][k])
            if i >= 2:
                main.wait_event(P["down"][k])                  # download of step i-2 has drained dout[k]
            P["graph"][k].replay()
            self.replays += 1
            P["ready"][k].record(main)
            with torch.cuda.stream(P["d2h"]):
                P["d2h"].wait_event(P["ready"][k])
                host_outputs[i].copy_(P["dout"][k], non_blocking=True)
                P["down"][k].record(P["d2h"])
        main.wait_stream(P["d2h"])
        P["d2h"].synchronize()                 # the pinned outputs are complete when this returns


# --------------------------------------------------------------------------- training (reference :44-46)
_TRAIN = {"flat": None, "last": None}


def begin_training():
    """Switch inference()/loss_image() to recording mode: the next inference() call keeps what the
    backward pass needs (the eager counterpart of building the gradient graph once)."""
    flow_architecture.TAPE = []


def end_training():
    flow_architecture.TAPE = None
    _TRAIN["last"] = None


def loss_image_train(sflow_p, sflow):
    """loss_image() for a recorded forward: remembers (prediction, target) for train()."""
    _TRAIN["last"] = (sflow_p, sflow)
    return loss_image(sflow_p, sflow)


def train(total_loss, lr):
    """reference :44-46 -- AdamOptimizer(lr).minimize(total_loss), executed eagerly: backward through the
    recorded forward, one flat fp32 gradient all-reduce when torch.distributed is initialised, TF1 Adam.

    Usage (mirrors train/flow_train.py:34-40,83):
        flow_net.begin_training()
        sflow_p = flow_net.inference(boundary, keep_prob)
        error = flow_net.loss_image_train(sflow_p, sflow)
        flow_net.train(error, learning_rate)
    Returns the loss tensor recomputed by the fused tanh+l2 backward kernel (fp32 [1])."""
    import model.flow_backward as fb
    tape = flow_architecture.TAPE
    if tape is None or not tape or _TRAIN["last"] is None:
        raise RuntimeError("train(): call begin_training(), inference() and loss_image_train() first")
    if _TRAIN["flat"] is None:
        _TRAIN["flat"] = fb.FlatParams()
        # the recorded forward used the pre-flattening tensors for its weights only; activations are unaffected
    sflow_p, sflow = _TRAIN["last"]
    loss = fb.backward(tape, sflow_p, sflow, _TRAIN["flat"])
    fb.adam_update(_TRAIN["flat"], lr)
    flow_architecture.TAPE = []          # ready for the next step
    _TRAIN["last"] = None
    return loss


class CompiledTraining:
    """One training step (forward + backward through the ~370 kernels of libflownet.so, incl. the re-packing of the
    weights Adam just changed) captured into CUDA graphs for a fixed (B, H, W) and replayed; the gradient all-reduce
    and the Adam update stay outside.  At small per-GPU batches -- BASELINE configs[2] sharded over 4-8 GPUs -- the eager
    step is bound by the host's launch rate, the replay is not.

    The step is TWO graphs (flow_backward.bucket_split): A = forward + backward of the decoder and the deepest level,
    after which the tail of the flat gradient (~86 % of the parameters) is final; B = backward of the encoder.  With
    torch.distributed initialised the first bucket's NCCL all-reduce is issued on a side stream behind A and runs under
    B; the second, small bucket follows B (bucket / overlap with backward).  Every rank reduces the same two
    slices in the same order, so the result is the single all-reduce's.

    With keep_prob < 1 the kernels add a device-resident counter to their dropout seeds (fn_conv_desc.dropout_seed_dev);
    step() bumps it before every replay, so each replay draws a new mask although the launch arguments are frozen."""

    def __init__(self, batch_size, shape=(128, 256), keep_prob=1.0, in_dtype=torch.uint8, dropout_seed=0):
        import model.flow_backward as fb
        self.fb = fb
        self.keep_prob, self.dropout_seed = keep_prob, dropout_seed
        self.seed_dev = torch.zeros(1, dtype=torch.int64, device="cuda") if keep_prob < 1.0 else None
        self.static_boundary = torch.zeros((batch_size, shape[0], shape[1], 1), dtype=in_dtype, device="cuda")
        self.static_sflow = torch.zeros((batch_size, shape[0], shape[1], 2), dtype=torch.float32, device="cuda")
        self._side_streams = os.environ.get("FLOWNET_TRAIN_STREAMS", "1") != "0"
        self._wstream = torch.cuda.Stream()
        begin_training()
        # warm-up: creates the variables, the flat buffers, the workspace and the allocator pool; gradients only (no
        # Adam), so the parameters are exactly what they were
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                self._part_a()
                self._part_b()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        flow_architecture.VARIABLES.touch()              # the capture must contain the weight-packing kernels
        self.graph = torch.cuda.CUDAGraph()              # part A
        self.graph_b = None
        n0 = native.launch_count()
        # inside the graphs: the weight images on one side stream (flow_architecture.prep_stream), the parameter gradients
        # on another (flow_backward.BackwardPass.wstream) -- the main chain is forward + data gradients only
        self._prep_keep = []
        prep = flow_architecture.prep_stream if self._side_streams else (lambda keep: contextlib.nullcontext())
        with torch.cuda.graph(self.graph):
            with prep(self._prep_keep):
                self._part_a()
        self.loss = self._pass.loss
        if self._cut is not None:
            self.graph_b = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_b, pool=self.graph.pool()):
                with prep(self._prep_keep):
                    self._part_b()
        self.launches_per_replay = native.launch_count() - n0
        self.bucket_offset = self._cut[1] if self._cut is not None else 0
        self._pass = None
        flow_architecture.TAPE = []
        self._comm = torch.cuda.Stream()
        self._ev = torch.cuda.Event()

    def _part_a(self):
        """forward, loss gradient, backward of tape[k:] (everything when the network has no cut point)"""
        flow_architecture.TAPE = []
        native.DROPOUT_SEED_DEV = self.seed_dev
        try:
            sflow_p = inference(self.static_boundary, self.keep_prob, dropout_seed=self.dropout_seed)
            flat = ensure_training_state()
            tape = flow_architecture.TAPE
            k, off = self.fb.bucket_split(tape, flat)
            self._cut = None if k is None else (k, off)
            self._tape = tape
            self._pass = self.fb.BackwardPass(sflow_p, self.static_sflow, flat)
            if self._side_streams and torch.cuda.is_current_stream_capturing():
                self._pass.wstream = self._wstream
            self._pass.run(tape if k is None else tape[k:])
            self._pass.join()
        finally:
            native.DROPOUT_SEED_DEV = None

    def _part_b(self):
        if self._cut is None:
            return
        native.DROPOUT_SEED_DEV = self.seed_dev
        try:
            self._pass.run(self._tape[:self._cut[0]])
            self._pass.join()
        finally:
            native.DROPOUT_SEED_DEV = None

    def step(self, boundary, sflow, lr):
        """boundary / sflow: CUDA or pinned host tensors of the captured shape.  Returns the loss (fp32 [1], device)."""
        import torch.distributed as dist
        flat = ensure_training_state()
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.static_boundary.copy_(boundary, non_blocking=True)
        self.static_sflow.copy_(sflow, non_blocking=True)
        if self.seed_dev is not None:
            self.seed_dev.add_(0x1E3779B97F4A7C15)           # new mask for this replay
        self.graph.replay()
        if self.graph_b is None:
            self.fb.adam_update(flat, lr)
            return self.loss
        if multi:
            # bucket 1 (decoder + deepest level) behind graph A on the side stream, under graph B
            main = torch.cuda.current_stream()
            self._ev.record(main)
            self._comm.wait_event(self._ev)
            with torch.cuda.stream(self._comm):
                dist.all_reduce(flat.grad[self.bucket_offset:], op=dist.ReduceOp.SUM)
        self.graph_b.replay()
        if multi:
            dist.all_reduce(flat.grad[:self.bucket_offset], op=dist.ReduceOp.SUM)
            torch.cuda.current_stream().wait_stream(self._comm)
        self.fb.adam_update(flat, lr, reduced=True)
        return self.loss


def boundary_gradient(boundary, sflow, keep_prob=1.0):
    """d l2_loss(inference(boundary), sflow) / d boundary -- the gradient the README's boundary-optimisation use needs
    (/root/reference/README.md:31-36): the training backward pass without the parameter gradients,
    plus the first conv's data gradient.  Returns (loss fp32 [1], gradient float32 [B,H,W,1])."""
    import model.flow_backward as fb
    saved = flow_architecture.TAPE
    flow_architecture.TAPE = []
    try:
        sflow_p = inference(boundary, keep_prob)
        loss, d_in = fb.backward(flow_architecture.TAPE, sflow_p, sflow, None, want_input_grad=True)
    finally:
        flow_architecture.TAPE = saved
    return loss, d_in.to(torch.float32)


def ensure_training_state():
    """Create the flat parameter / gradient / Adam buffers now (needs one forward: variables are created lazily)."""
    import model.flow_backward as fb
    if _TRAIN["flat"] is None:
        _TRAIN["flat"] = fb.FlatParams()
    return _TRAIN["flat"]


def training_state():
    """The flat parameter / gradient / Adam buffers (None before the first train() call)."""
    return _TRAIN["flat"]


def reset_training_state():
    _TRAIN["flat"] = None
    _TRAIN["last"] = None
