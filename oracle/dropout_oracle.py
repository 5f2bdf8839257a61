"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the counter-based dropout mask of libflownet.so.

tf.nn.dropout (TF1; /root/reference/model/flow_architecture.py:144-145, keep_prob 0.7 in train/flow_train.py:21) draws
`floor(keep_prob + U[0,1))` from TF's stateful RNG, which cannot be reproduced outside TensorFlow.  The CUDA path uses
a counter-based generator instead (csrc/common.cuh: mix64 / dropout_scale) so that the forward transform, the
recomputation in the backward pass, the weight-gradient transform and the prologue adjoint all see the SAME mask
without storing it.  This file restates that generator bit for bit, so the oracle's autograd can run a whole
keep_prob < 1 training step under the identical mask (tests/test_gpu_parity.py, tests/test_oracle.py).

    element idx of the dropped tensor [B,H,W,K] (flat NHWC index, K = channels AFTER the nonlinearity):
      q = idx >> 3;  h1 = mix64(seed ^ mix64(q));  h2 = mix64(h1)
      u = 16-bit field (idx & 3) of (idx & 4 ? h2 : h1)
      keep iff u < uint32(float32(keep_prob) * 65536);  kept values are scaled by float32(1 / keep_prob)
    seed of res_block number n (1-based, in call order of conv_res) = dropout_seed + n
"""
import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def mix64(z):
    """splitmix64 finaliser on uint64 arrays (wrap-around arithmetic)"""
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def threshold(keep_prob):
    return np.uint32(np.float32(keep_prob) * np.float32(65536.0))


def dropout_scale(seed, shape, keep_prob):
    """float32 array of `shape` ([B,H,W,K]): 1/keep_prob where the element is kept, 0 where it is dropped."""
    n = int(np.prod(shape))
    idx = np.arange(n, dtype=np.uint64)
    seed = np.uint64(int(seed) & 0xFFFFFFFFFFFFFFFF)
    h1 = mix64(seed ^ mix64(idx >> np.uint64(3)))
    h2 = mix64(h1)
    h = np.where((idx & np.uint64(4)) != 0, h2, h1)
    u = ((h >> (np.uint64(16) * (idx & np.uint64(3)))) & np.uint64(0xFFFF)).astype(np.uint32)
    inv = np.float32(1.0) / np.float32(keep_prob)
    return np.where(u < threshold(keep_prob), inv, np.float32(0.0)).astype(np.float32).reshape(shape)
