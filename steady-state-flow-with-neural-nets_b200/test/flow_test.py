"""The reference's qualitative test (/root/reference/test/flow_test.py:52-100) on the B200 path: restore the newest
checkpoint of the flag-named directory, run `flow_net.inference(boundary, 1.0)` one image at a time over the test
set, and build the reference's display image `[true | generated | true - generated]` (speed magnitude minus
0.05 * boundary, :89-92).  matplotlib is not part of this image, so instead of `plt.show()` the display images are
written as .npy (and .png when cv2 is importable) under --out_dir and the per-image errors are printed.

Test set: `../data/computed_car_flow/*/fluid_flow_0002.h5` as in the reference (needs h5py), else the boundaries /
targets of `--data_glob` TFRecords, else the bundled car masks with no ground truth (generated field only).

    python flow_test.py [--base_dir=../checkpoints] [--data_glob='../data/*.tfrecords'] [--out_dir=./flow_test_out]
"""
import glob
import os
import re
import sys

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
flags.DEFINE_bool('display_test', True, """ display the test images """)
flags.DEFINE_string('data_glob', '../data/*.tfrecords', """ TFRecords to test on when no computed_car_flow directory exists """)
flags.DEFINE_string('out_dir', './flow_test_out', """ where the display images go (no matplotlib here) """)


def tryint(s):
    try:
        return int(s)
    except ValueError:
        return s


def alphanum_key(s):
    """reference :46-47"""
    return [tryint(c) for c in re.split('([0-9]+)', s)]


def display_image(boundary, sflow_true, sflow_generated):
    """reference :88-92: [true | generated | difference] speed magnitude, solids darkened."""
    sflow_plot = np.concatenate([sflow_true, sflow_generated, sflow_true - sflow_generated], axis=1)
    boundary_concat = np.concatenate(3 * [boundary], axis=1)
    return np.sqrt(np.square(sflow_plot[:, :, 0]) + np.square(sflow_plot[:, :, 1])) - .05 * boundary_concat[:, :, 0]


def test_set(shape):
    """Yields (name, boundary [H,W,1] float32, sflow_true [H,W,2] float32 or None)."""
    runs = sorted(glob.glob('../data/computed_car_flow/*/'), key=alphanum_key)
    if runs:
        from utils.flow_reader import load_boundary, load_flow
        for run in runs:
            name = run + '/fluid_flow_0002.h5'
            yield run, np.float32(load_boundary(name, shape)), np.float32(load_flow(name, shape))
        return
    files = sorted(glob.glob(FLAGS.data_glob))
    if files:
        from input.flow_input import read_data
        for f in files:
            for i, (b, s) in enumerate(read_data(f, tuple(shape))):
                yield '%s[%d]' % (f, i), b.astype(np.float32), s
        return
    cars = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..', 'tests', 'golden', 'cars_128x256.npz')
    if os.path.exists(cars):
        from utils.synthetic import load_car_boundaries
        for i, b in enumerate(load_car_boundaries(cars)):
            yield 'car_%03d' % (i + 1), b.astype(np.float32).reshape(shape[0], shape[1], 1), None


def evaluate(max_images=None):
    """Run Eval once (reference :52-96).  Returns [(name, rel-L2 error or None)]."""
    shape = [128, 256]
    test_dir = make_checkpoint_path(FLAGS.base_dir, FLAGS)
    ckpt = checkpoint.get_checkpoint_state(test_dir)
    if ckpt is None:
        print("no checkpoint under " + test_dir + ": testing the seeded random initialisation")
        flow_architecture.reset_variables(1234)
    else:
        flow_architecture.reset_variables(1234)
        checkpoint.restore(ckpt, flow_architecture.VARIABLES, strict=False)
    results = []
    for k, (name, boundary_np, sflow_true) in enumerate(test_set(shape)):
        if max_images is not None and k >= max_images:
            break
        sflow_generated = flow_net.inference(torch.from_numpy(boundary_np[None]).cuda(), 1.0)[0].cpu().numpy()
        err = None
        if sflow_true is not None:
            err = float(np.linalg.norm(sflow_true - sflow_generated) / max(np.linalg.norm(sflow_true), 1e-30))
        results.append((name, err))
        print(name, "rel-L2 error", err)
        if FLAGS.display_test:
            os.makedirs(FLAGS.out_dir, exist_ok=True)
            img = display_image(boundary_np, sflow_true if sflow_true is not None else np.zeros_like(sflow_generated),
                                sflow_generated)
            base = os.path.join(FLAGS.out_dir, 'flow_%04d' % k)
            np.save(base + '.npy', img)
            try:
                import cv2
                lo, hi = float(img.min()), float(img.max())
                cv2.imwrite(base + '.png', np.uint8(255 * (img - lo) / max(hi - lo, 1e-12)))
            except ImportError:
                pass
    return results


def main(argv=None):
    evaluate()


if __name__ == '__main__':
    main()
