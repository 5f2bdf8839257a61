"""Training loop with the reference's flags and shape (/root/reference/train/flow_train.py:15-24, 30-97):
make inputs, unwrap the network, l2 loss, Adam step; print the loss every 100 steps; NaN check.
The TFRecord queue is input/flow_input.py's (native parser, shuffle buffer) when `--data_glob` matches files, else
the synthetic generator (no dataset ships, no network); the TF
Saver by utils/checkpoint.py: an .npz keyed by the TF variable names (+ Adam slots) every 1000 steps and a
restore-on-start with the reference's fall-back to random init.
Data-parallel: launch with torchrun, one process per GPU; each rank trains on its own shard of the
batch and the flat fp32 gradient is summed with one NCCL all-reduce per step.

    python flow_train.py [--batch_size=8] [--max_steps=200] [--keep_prob=0.7] [--learning_rate=1e-4]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.append(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))

import model.flow_architecture as flow_architecture  # noqa: E402
import model.flow_net as flow_net  # noqa: E402
from model import flags  # noqa: E402
from utils import checkpoint  # noqa: E402
from utils.experiment_manager import make_checkpoint_path  # noqa: E402

FLAGS = flags.FLAGS
flags.DEFINE_string('base_dir', '../checkpoints', """dir to store trained net """)
flags.DEFINE_integer('batch_size', 8, """ training batch size """)
flags.DEFINE_integer('max_steps', 500000, """ max number of steps to train """)
flags.DEFINE_float('keep_prob', 0.7, """ keep probability for dropout """)
flags.DEFINE_float('learning_rate', 1e-4, """ keep probability for dropout """)
flags.DEFINE_bool('cuda_graph', True, """ replay forward + backward from one CUDA graph (flow_net.CompiledTraining) """)
flags.DEFINE_string('data_glob', '../data/*.tfrecords', """ dataset files (reference: input/flow_input.py:45); synthetic data if none match """)


def train():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl")
    train_dir = make_checkpoint_path(FLAGS.base_dir, FLAGS)
    if rank == 0:
        print(train_dir)
    per_rank = max(1, FLAGS.batch_size // world)
    import glob
    data_glob = FLAGS.data_glob if glob.glob(FLAGS.data_glob) else None
    if rank == 0:
        print("training on " + (FLAGS.data_glob if data_glob else "synthetic boundaries (no dataset found)"))

    def build(restored):
        """Variables (every rank starts from the same ones), the captured step, and -- reference :61-72 -- the checkpoint.
        The variables are created lazily, so a checkpoint can only be validated against the graph once a forward (the
        captured step's warm-up, or one eager inference) has run: everything that can raise ValueError happens in here,
        and the caller falls back to the random init."""
        flow_architecture.reset_variables(1234)
        flow_net.reset_training_state()
        start = 0
        if restored is not None:
            start = checkpoint.restore(restored, flow_architecture.VARIABLES) + 1      # initial values of the lazy variables
        flow_net.begin_training()
        comp = None
        if FLAGS.cuda_graph:
            # captures one step for this (batch, shape); its warm-up creates the variables and the flat buffers without
            # changing the parameters
            comp = flow_net.CompiledTraining(per_rank, (128, 256), keep_prob=FLAGS.keep_prob, dropout_seed=1 + rank)
        else:
            b0, _ = flow_net.inputs(per_rank, seed=0)
            flow_net.inference(b0, 1.0)
            flow_architecture.TAPE = []
        if restored is not None:
            # now that the variables exist: shapes and names are checked (strict), the Adam slots restored
            checkpoint.restore(restored, flow_architecture.VARIABLES, flow_net.ensure_training_state(), strict=True)
        return comp, start

    # init from checkpoint (reference :61-72): every rank reads the same file; a checkpoint that does not fit the graph is
    # discarded and training starts from the random init
    ckpt = checkpoint.get_checkpoint_state(train_dir)
    restored = None
    if ckpt is not None:
        if rank == 0:
            print("init from " + train_dir)
        restored = checkpoint.load(ckpt)
    try:
        compiled, start_step = build(restored)
    except ValueError:
        if restored is None:
            raise
        print("there was a problem using variables in checkpoint, random init will be used instead")
        compiled, start_step = build(None)
    for step in range(start_step, FLAGS.max_steps):
        t = time.time()
        boundary, sflow = flow_net.inputs(per_rank, seed=(step * world + rank) * per_rank if not data_glob else rank,
                                          data_glob=data_glob, shard=(rank, world))
        if compiled is not None:
            error = compiled.step(boundary, sflow, FLAGS.learning_rate)
        else:
            sflow_p = flow_net.inference(boundary, FLAGS.keep_prob, dropout_seed=step * 131071 + rank)
            error = flow_net.loss_image_train(sflow_p, sflow)
            flow_net.train(error, FLAGS.learning_rate)
        loss_value = float(error.item())
        elapsed = time.time() - t
        assert not np.isnan(loss_value), 'Model diverged with loss = NaN'
        if step % 100 == 0 and rank == 0:
            print("loss value at " + str(loss_value))
            print("time per batch is " + str(elapsed))
        if step % 1000 == 0 and rank == 0 and step > 0:
            checkpoint.save(train_dir, step, flow_architecture.global_variables(), flow_net.training_state())
            print("saved to " + train_dir)
    flow_net.end_training()


def main(argv=None):
    train()


if __name__ == '__main__':
    main()
