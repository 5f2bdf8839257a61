"""Dataset writer (/root/reference/utils/createTFRecords.py): one tf.train.Example per simulation, features
'boundary' (uint8 [128*256]) and 'sflow' (float32 [128*256*2]) in a TFRecord file.  Sources: MechSys runs
(`<dir>/*/fluid_flow_0002.h5`, needs h5py) or, without a dataset, the synthetic generator.

    python createTFRecords.py --out ../data/train.tfrecords [--h5_glob '/data/fluid_flow_steady_state_128x128_/*'] [--count 64]
"""
import argparse
import glob
import os
import sys

import numpy as np

sys.path.append(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))


def main():
    from input.flow_input import write_tfrecords
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default='../data/train.tfrecords')
    ap.add_argument('--h5_glob', default=None)
    ap.add_argument('--count', type=int, default=64, help='synthetic examples when no --h5_glob is given')
    ap.add_argument('--seed', type=int, default=0)
    a = ap.parse_args()
    shape = [128, 256]
    if a.h5_glob:
        from utils.flow_reader import load_boundary, load_flow
        runs = sorted(glob.glob(a.h5_glob))
        b = np.stack([np.uint8(load_boundary(r + '/fluid_flow_0002.h5', shape)) for r in runs])
        s = np.stack([np.float32(load_flow(r + '/fluid_flow_0002.h5', shape)) for r in runs])
    else:
        from utils.synthetic import synth_boundary, synth_sflow
        b = synth_boundary(a.count, shape[0], shape[1], seed=a.seed)
        s = synth_sflow(b)
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    write_tfrecords(a.out, b, s)
    print("wrote %d examples to %s" % (len(b), a.out))


if __name__ == '__main__':
    main()
