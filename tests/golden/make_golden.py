#!/usr/bin/env python
"""Generates the committed fixtures under tests/golden/.  Run in the BUILD container (it reads
/root/reference/cars/*.txt, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Fixtures
  cars_128x256.npz   28 car boundary masks (bit-packed uint8) placed/cropped as the reference's
                     LBM generator + flow_reader do (oracle.flow_oracle.car_boundary_from_txt)
  golden_tiny.npz    fp64 oracle forward of one seeded 16x32 and one 32x48 boundary through the
                     default-flag net (seed-1234 Xavier variables), plus every res-block output
  golden_cars.npz    fp64->fp32 oracle forward of cars 1 and 15 at 128x256 (default flags)
  golden_vars.npz    fingerprints (sum, abs-sum, first 4 values) of every seeded variable, which
                     pin the RNG draw order shared by the oracle and the product's variable store
  golden_train.npz   loss, a few gradient fingerprints and the post-Adam value of two variables
                     for one 2-image 32x64 training step (fp64)
The TensorFlow reference cannot run here, so these pin the ORACLE, not TF.
"""
import glob
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import flow_oracle as fo  # noqa: E402


def main():
    torch.set_num_threads(8)
    # ---- cars
    paths = sorted(glob.glob("/root/reference/cars/car_*.txt"))
    assert len(paths) == 28, len(paths)
    cars = np.stack([fo.car_boundary_from_txt(p) for p in paths]).astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "cars_128x256.npz"),
                        packed=np.packbits(cars, axis=None), shape=np.array(cars.shape))

    # ---- variable fingerprints
    vs = fo.VariableStore(1234, torch.float64)
    x = torch.from_numpy(fo.synth_boundary(1, 16, 32, seed=0)).double()
    taps = {}
    y = fo.inference(vs, x, 1.0, taps=taps)
    names = list(vs.vars.keys())
    fp = np.array([[v.sum().item(), v.abs().sum().item()] + v.flatten()[:4].tolist() +
                   [0.0] * max(0, 4 - v.numel()) for v in vs.vars.values()])
    np.savez_compressed(os.path.join(HERE, "golden_vars.npz"), names=np.array(names), fp=fp,
                        shapes=np.array([str(list(v.shape)) for v in vs.vars.values()]))

    # ---- tiny forward
    x2 = torch.from_numpy(fo.synth_boundary(2, 32, 48, seed=5)).double()
    y2 = fo.inference(vs, x2, 1.0)
    np.savez_compressed(os.path.join(HERE, "golden_tiny.npz"),
                        x=x.numpy().astype(np.uint8), y=y.numpy(),
                        x2=x2.numpy().astype(np.uint8), y2=y2.numpy(),
                        **{"tap_" + k: v.numpy() for k, v in taps.items()})

    # ---- cars forward (two of them)
    idx = [0, 14]
    xc = torch.from_numpy(cars[idx].astype(np.float64))[..., None]
    yc = fo.inference(vs, xc, 1.0)
    np.savez_compressed(os.path.join(HERE, "golden_cars.npz"), idx=np.array(idx),
                        y=yc.numpy().astype(np.float32))

    # ---- one training step
    vt = fo.VariableStore(1234, torch.float64)
    xb = torch.from_numpy(fo.synth_boundary(2, 32, 64, seed=3)).double()
    rng = np.random.default_rng(99)
    tgt = torch.from_numpy(0.1 * np.tanh(rng.standard_normal((2, 32, 64, 2)))).double() * (1 - xb)
    fo.inference(vt, xb, 1.0)          # create variables
    params = {k: v.clone().requires_grad_(True) for k, v in vt.vars.items()}
    vt.vars = params
    loss = fo.loss_image(fo.inference(vt, xb, 1.0), tgt)
    loss.backward()
    grads = {k: p.grad.clone() for k, p in params.items()}
    pd = {k: p.detach().clone() for k, p in params.items()}
    m = {k: torch.zeros_like(p) for k, p in pd.items()}
    v = {k: torch.zeros_like(p) for k, p in pd.items()}
    fo.adam_step_tf1(pd, grads, m, v, t=1)
    gnames = list(grads.keys())
    gfp = np.array([[g.sum().item(), g.abs().sum().item(), g.pow(2).sum().sqrt().item()]
                    for g in grads.values()])
    np.savez_compressed(os.path.join(HERE, "golden_train.npz"), loss=loss.item(),
                        names=np.array(gnames), gfp=gfp,
                        g_last_w=grads["last_conv_conv/weights"].numpy(),
                        g_first_w=grads["resnet_1_0_conv_1_conv/weights"].numpy(),
                        p_last_w=pd["last_conv_conv/weights"].numpy(),
                        p_first_b=pd["resnet_1_0_conv_1_conv/biases"].numpy(),
                        target=tgt.numpy())
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
