"""MechSys .h5 crop rule of the reference (/root/reference/utils/flow_reader.py:5-17): the simulator writes
`Velocity_0` ([H, W+128, 3] flattened) and `Gamma` ([H, W+128, 1]); the dataset keeps the first W columns and the
x/y velocity components.  h5py is not part of this image: the crop itself is plain numpy (tested), the file access
raises a clear error when h5py is missing."""
import numpy as np


def crop_flow(velocity_flat, shape):
    return np.asarray(velocity_flat).reshape([shape[0], shape[1] + 128, 3])[0:shape[0], 0:shape[1], 0:2]


def crop_boundary(gamma_flat, shape):
    return np.asarray(gamma_flat).reshape([shape[0], shape[1] + 128, 1])[0:shape[0], 0:shape[1], :]


def _h5():
    try:
        import h5py
    except ImportError as e:                                  # pragma: no cover (h5py absent in this image)
        raise ImportError("reading MechSys .h5 files needs h5py, which is not installed") from e
    return h5py


def load_flow(filename, shape):
    """reference :5-10"""
    with _h5().File(filename, 'r') as f:
        return crop_flow(np.array(f['Velocity_0'][:]), shape)


def load_boundary(filename, shape):
    """reference :12-17"""
    with _h5().File(filename, 'r') as f:
        return crop_boundary(np.array(f['Gamma'][:]), shape)
