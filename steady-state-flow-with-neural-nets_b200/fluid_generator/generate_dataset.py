"""Generate a labelled dataset on the box: random obstacle boundaries by the reference's placement rules
(generate_xiao_fluid_flow.cpp:262-340, utils/synthetic.py), flows by the GPU lattice-Boltzmann solver, written as the
TFRecords the training script reads (utils/createTFRecords.py:45-56).

    python generate_dataset.py --count 256 --out ../data/train.tfrecords [--steps 10000] [--batch 128] [--fp64]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.append(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))


def main():
    from fluid_generator import lbm
    from input.flow_input import write_tfrecords
    from utils.synthetic import synth_boundary
    ap = argparse.ArgumentParser()
    ap.add_argument('--count', type=int, default=64)
    ap.add_argument('--batch', type=int, default=128)
    ap.add_argument('--steps', type=int, default=lbm.STEPS)
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--fp64', action='store_true')
    ap.add_argument('--out', default='../data/train.tfrecords')
    a = ap.parse_args()
    bs, ss = [], []
    for i0 in range(0, a.count, a.batch):
        n = min(a.batch, a.count - i0)
        boundary = synth_boundary(n, 128, 256, seed=a.seed + i0)
        b, s = lbm.simulate(boundary, steps=a.steps, dtype=torch.float64 if a.fp64 else torch.float32)
        bs.append(b)
        ss.append(s.cpu().numpy())
        print("simulated %d / %d" % (i0 + n, a.count), flush=True)
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    write_tfrecords(a.out, np.concatenate(bs), np.concatenate(ss))
    print("wrote %d examples to %s" % (a.count, a.out))


if __name__ == '__main__':
    main()
