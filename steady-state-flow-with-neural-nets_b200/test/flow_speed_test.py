"""The reference's speed test (/root/reference/test/flow_speed_test.py:50-93) on the B200 path:
batch-8 all-zero boundaries at 128x256, 1 warm-up + num_runs timed inference calls with host feed
and host fetch every call, prints the time per image.  The reference restores a trained checkpoint
first; none ships (and timing does not depend on the values), so seeded Xavier weights are used.

    python flow_speed_test.py [--batch_size=8] [--num_runs=1000]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.append(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))

import model.flow_net as flow_net  # noqa: E402
from model import flags  # noqa: E402

FLAGS = flags.FLAGS
flags.DEFINE_string('base_dir', '../checkpoints', """dir to store trained net """)
flags.DEFINE_integer('batch_size', 8, """ training batch size """)
flags.DEFINE_integer('max_steps', 500000, """ max number of steps to train """)
flags.DEFINE_float('keep_prob', 0.7, """ keep probability for dropout """)
flags.DEFINE_float('learning_rate', 1e-4, """ keep probability for dropout """)
flags.DEFINE_integer('num_runs', 1000, """ timed inference calls (reference: 1000) """)


def evaluate():
    shape = [128, 256]
    B = FLAGS.batch_size
    run = flow_net.CompiledInference(B, tuple(shape), keep_prob=1.0, in_dtype=torch.float32)
    boundary_np = torch.zeros([B, shape[0], shape[1], 1]).pin_memory()        # feed_dict source
    out_np = torch.empty([B, shape[0], shape[1], 2]).pin_memory()             # fetch target
    out_np.copy_(run(boundary_np))
    torch.cuda.synchronize()
    t = time.time()
    for _ in range(FLAGS.num_runs):
        out_np.copy_(run(boundary_np), non_blocking=True)
        torch.cuda.synchronize()                 # sess.run returns host data: one sync per call
    elapsed = time.time() - t
    print("time per input is ")
    print(elapsed / (FLAGS.num_runs * float(B)))
    assert np.isfinite(out_np.numpy()).all()


def main(argv=None):
    evaluate()


if __name__ == '__main__':
    main()
