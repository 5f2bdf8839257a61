"""CPU tests that pin the oracle: against the committed golden vectors, against the independent
naive-NumPy restatement, and through the adjointness properties that define TF's conv2d_transpose.
(The TF reference itself cannot run here -- parity is unpinned at the TF boundary)"""
import os

import numpy as np
import pytest
import torch

from oracle import flow_oracle as fo
from oracle import naive_numpy as nn_


def test_variable_inventory_matches_reference_layer_table():
    vs = fo.VariableStore(1234)
    x = torch.zeros(1, 16, 32, 1, dtype=torch.float64)
    fo.inference(vs, x, 1.0)
    assert len(vs.vars) == 70                       # 35 weight layers
    assert vs.num_params() == 2_561_386
    names = list(vs.vars)
    assert names[0] == "resnet_1_0_conv_1_conv/weights" and names[1] == "resnet_1_0_conv_1_conv/biases"
    assert names[-2] == "last_conv_conv/weights"
    assert list(vs.vars["up_conv_1_trans_conv/weights"].shape) == [3, 3, 64, 128]
    assert list(vs.vars["resnet_up_1_0_nin_fc/weights"].shape) == [128, 64]
    # creation order inside a skip block: conv_1, nin, conv_2 (flow_architecture.py:132-149)
    i = names.index("resnet_up_1_0_conv_1_conv/weights")
    assert names[i + 2] == "resnet_up_1_0_nin_fc/weights" and names[i + 4] == "resnet_up_1_0_conv_2_conv/weights"


def test_config5_inventory():
    vs = fo.VariableStore(1234)
    fo.inference(vs, torch.zeros(1, 16, 32, 1, dtype=torch.float64), 1.0, nr_res_blocks=2, filter_size=16)
    assert vs.num_params() == 16_132_210


def test_golden_variable_fingerprints(golden_dir):
    z = np.load(os.path.join(golden_dir, "golden_vars.npz"))
    vs = fo.VariableStore(1234)
    fo.inference(vs, torch.zeros(1, 16, 32, 1, dtype=torch.float64), 1.0)
    assert list(z["names"]) == list(vs.vars)
    for row, v in zip(z["fp"], vs.vars.values()):
        assert abs(row[0] - v.sum().item()) < 1e-9
        assert abs(row[1] - v.abs().sum().item()) < 1e-9


def test_golden_tiny_forward(golden_dir):
    z = np.load(os.path.join(golden_dir, "golden_tiny.npz"))
    vs = fo.VariableStore(1234)
    taps = {}
    y = fo.inference(vs, torch.from_numpy(z["x"].astype(np.float64)), 1.0, taps=taps)
    np.testing.assert_allclose(y.numpy(), z["y"], rtol=0, atol=1e-12)
    for k, v in taps.items():
        np.testing.assert_allclose(v.numpy(), z["tap_" + k], rtol=0, atol=1e-11)
    y2 = fo.inference(vs, torch.from_numpy(z["x2"].astype(np.float64)), 1.0)
    np.testing.assert_allclose(y2.numpy(), z["y2"], rtol=0, atol=1e-12)


def test_golden_inputs_are_the_synthetic_generator(golden_dir):
    z = np.load(os.path.join(golden_dir, "golden_tiny.npz"))
    np.testing.assert_array_equal(z["x"], fo.synth_boundary(1, 16, 32, seed=0).astype(np.uint8))
    np.testing.assert_array_equal(z["x2"], fo.synth_boundary(2, 32, 48, seed=5).astype(np.uint8))


@pytest.mark.parametrize("shape", [(1, 16, 32), (2, 20, 24), (1, 17, 33)])
def test_oracle_vs_naive_numpy(shape):
    B, H, W = shape
    vs = fo.VariableStore(7)
    rng = np.random.default_rng(3)
    x = (rng.random((B, H, W, 1)) > 0.6).astype(np.float64)
    y = fo.inference(vs, torch.from_numpy(x), 1.0).numpy()
    yn = nn_.conv_res({k: v.numpy() for k, v in vs.vars.items()}, x)
    assert y.shape == yn.shape
    np.testing.assert_allclose(y, yn, rtol=0, atol=1e-12)


@pytest.mark.parametrize("H,W,s", [(8, 10, 1), (8, 10, 2), (7, 9, 2), (9, 6, 1)])
def test_conv_same_matches_naive(H, W, s):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, H, W, 3))
    w = rng.standard_normal((3, 3, 3, 4))
    b = rng.standard_normal(4)
    y = (fo.conv2d_same(torch.from_numpy(x), torch.from_numpy(w), s) + torch.from_numpy(b)).numpy()
    np.testing.assert_allclose(y, nn_.conv2d_same(x, w, b, s), atol=1e-12)


@pytest.mark.parametrize("h,w", [(4, 5), (3, 8)])
def test_conv_transpose_is_adjoint_of_strided_same_conv(h, w):
    """<conv_s2(X), Y> == <X, conv_transpose(Y)> with the same [k,k,Cout_t,Cin_t] weights (App. B.2)."""
    rng = np.random.default_rng(1)
    cin_t, cout_t = 5, 3                        # transpose conv: cin_t -> cout_t, 2x upsampling
    wt = torch.from_numpy(rng.standard_normal((3, 3, cout_t, cin_t)))
    X = torch.from_numpy(rng.standard_normal((2, 2 * h, 2 * w, cout_t)))
    Y = torch.from_numpy(rng.standard_normal((2, h, w, cin_t)))
    lhs = (fo.conv2d_same(X, wt, 2) * Y).sum()          # wt read as HWIO with I=cout_t, O=cin_t
    rhs = (X * fo.conv2d_transpose_same(Y, wt, 2)).sum()
    assert abs(lhs.item() - rhs.item()) < 1e-9 * max(1.0, abs(lhs.item()))
    zn = nn_.conv2d_transpose_same(Y.numpy(), wt.numpy(), np.zeros(cout_t), 2)
    np.testing.assert_allclose(fo.conv2d_transpose_same(Y, wt, 2).numpy(), zn, atol=1e-12)


def test_strided_same_pads_bottom_right_only():
    """F4: k=3, s=2 on an even axis pads (0,1); torch conv2d(padding=1, stride=2) is NOT equivalent."""
    x = torch.arange(16, dtype=torch.float64).reshape(1, 4, 4, 1)
    w = torch.zeros(3, 3, 1, 1, dtype=torch.float64)
    w[0, 0, 0, 0] = 1.0                                   # picks x[2i, 2j] iff pad-before is 0
    y = fo.conv2d_same(x, w, 2)
    np.testing.assert_array_equal(y[0, :, :, 0].numpy(), x[0, ::2, ::2, 0].numpy())


def test_avg_pool_same_odd_edges():
    x = np.arange(2 * 5 * 3 * 2, dtype=np.float64).reshape(2, 5, 3, 2)
    np.testing.assert_allclose(fo.avg_pool_2x2_same(torch.from_numpy(x)).numpy(), nn_.avg_pool_2x2_same(x))


def test_front_channel_pad_and_gate():
    """F6: the shortcut occupies the LAST in_filter channels of a widening block."""
    vs = fo.VariableStore(0)
    x = torch.randn(1, 4, 4, 4, dtype=torch.float64)
    fo.res_block(vs, x, filter_size=8, gated=True, name="b")
    for k in vs.vars:
        vs.vars[k].zero_()                                # conv outputs 0 -> x_2 = 0*sigmoid(0) = 0
    y = fo.res_block(vs, x, filter_size=8, gated=True, name="b")
    assert torch.equal(y[..., :4], torch.zeros_like(y[..., :4]))
    assert torch.equal(y[..., 4:], x)


def test_l2_loss_and_tf1_adam():
    p = torch.tensor([1.0, -2.0, 3.0], dtype=torch.float64)
    t = torch.tensor([0.5, 0.0, 1.0], dtype=torch.float64)
    assert abs(fo.loss_image(p, t).item() - (0.25 + 4 + 4) / 2) < 1e-15
    params = {"a": torch.tensor([1.0], dtype=torch.float64)}
    g = {"a": torch.tensor([0.3], dtype=torch.float64)}
    m = {"a": torch.zeros(1, dtype=torch.float64)}
    v = {"a": torch.zeros(1, dtype=torch.float64)}
    fo.adam_step_tf1(params, g, m, v, t=1, lr=1e-4)
    # first step: m=0.1g*... -> update = lr*sqrt(1-b2)/(1-b1) * (0.1g)/(sqrt(0.001 g^2)+eps)
    lr_t = 1e-4 * np.sqrt(1 - 0.999) / (1 - 0.9)
    exp = 1.0 - lr_t * (0.1 * 0.3) / (np.sqrt(0.001 * 0.09) + 1e-8)
    assert abs(params["a"].item() - exp) < 1e-15


def test_golden_train_step(golden_dir):
    z = np.load(os.path.join(golden_dir, "golden_train.npz"))
    vt = fo.VariableStore(1234)
    xb = torch.from_numpy(fo.synth_boundary(2, 32, 64, seed=3)).double()
    tgt = torch.from_numpy(z["target"])
    fo.inference(vt, xb, 1.0)
    vt.vars = {k: v.clone().requires_grad_(True) for k, v in vt.vars.items()}
    loss = fo.loss_image(fo.inference(vt, xb, 1.0), tgt)
    loss.backward()
    assert abs(loss.item() - float(z["loss"])) < 1e-9 * abs(float(z["loss"]))
    np.testing.assert_allclose(vt.vars["last_conv_conv/weights"].grad.numpy(), z["g_last_w"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(vt.vars["resnet_1_0_conv_1_conv/weights"].grad.numpy(), z["g_first_w"], rtol=1e-9,
                               atol=1e-12)


def test_cars_fixture_shape_and_walls(golden_dir):
    from utils.synthetic import load_car_boundaries
    cars = load_car_boundaries(os.path.join(golden_dir, "cars_128x256.npz"))
    assert cars.shape == (28, 128, 256, 1) and cars.dtype == np.uint8
    assert cars[:, 0].all() and cars[:, 127].all()             # channel walls
    assert not cars[:, 1:28, :, :].any()                      # nothing above the car box
    frac = cars[:, 28:128, 32:256].mean()
    assert 0.2 < frac < 0.6
    if os.path.exists("/root/reference/cars/car_001.txt"):     # build container only
        ref = fo.car_boundary_from_txt("/root/reference/cars/car_001.txt")
        np.testing.assert_array_equal(cars[0, :, :, 0], ref.astype(np.uint8))


def test_golden_cars_forward(golden_dir):
    from utils.synthetic import load_car_boundaries
    z = np.load(os.path.join(golden_dir, "golden_cars.npz"))
    cars = load_car_boundaries(os.path.join(golden_dir, "cars_128x256.npz"))
    vs = fo.VariableStore(1234, torch.float32)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    y = fo.inference(vs, torch.from_numpy(cars[list(z["idx"])].astype(np.float32)), 1.0)
    err = (y - torch.from_numpy(z["y"])).norm() / torch.from_numpy(z["y"]).norm()
    assert err < 1e-5            # fp32 oracle vs stored fp64 oracle (BASELINE.md tolerance)


# ---------------------------------------------------------------- lattice-Boltzmann data generator
def test_lbm_oracle_conserves_mass_in_a_closed_periodic_box_and_reaches_poiseuille_flow():
    from oracle import lbm_oracle as lo
    # (1) collision + periodic streaming without boundary columns conserve mass and momentum exactly
    rng = np.random.default_rng(0)
    B, ny, nx = 1, 12, 16
    solid = np.zeros((B, ny, nx), bool)
    f = lo.init(B, ny, nx) * (1 + 0.05 * rng.standard_normal((9, B, ny, nx)))
    m0, px0 = f.sum(), (lo.CX[:, None, None, None] * f).sum()
    g = f.copy()
    for _ in range(5):
        rho, vx, vy = lo.macroscopic(g, solid)
        post = g - (g - lo.feq(rho, vx, vy)) / 0.65
        g = np.stack([np.roll(np.roll(post[k], lo.CY[k], axis=1), lo.CX[k], axis=2) for k in range(9)])
    assert abs(g.sum() - m0) < 1e-10 * m0 and abs((lo.CX[:, None, None, None] * g).sum() - px0) < 1e-10 * abs(px0)
    # (2) bounce-back solids hold mass too
    solid[:, 0, :] = solid[:, -1, :] = True
    post = np.where(solid[None], f[lo.OPP], f)
    assert abs(post.sum() - f.sum()) < 1e-12 * f.sum()
    # (3) an empty channel converges to the parabolic profile the inlet prescribes (:239-246)
    ny, nx = 24, 64
    solid = np.zeros((1, ny, nx), bool)
    solid[:, 0, :] = solid[:, -1, :] = True
    rho, vx, vy = lo.run(solid, 1500)
    mid = vx[0, :, nx // 2]
    # developed flow: symmetric parabola between the half-way bounce-back walls, carrying the inlet's flux.  (The inlet
    # formula itself, yp = i - 1.5, is off-centre by one cell -- kept as the reference has it.)
    assert np.abs(mid - mid[::-1]).max() < 2e-3
    yy = np.arange(ny) - 0.5
    shape = yy * (ny - 2 - yy)
    fit = mid[1:-1] @ shape[1:-1] / (shape[1:-1] @ shape[1:-1]) * shape
    assert np.abs(mid[1:-1] - fit[1:-1]).max() < 0.02 * mid.max()
    flux = (rho * vx)[0].sum(axis=0)
    assert abs(flux[nx // 4] - flux[3 * nx // 4]) < 3e-2 * flux[nx // 4]        # still settling after 1500 steps
    assert 0.09 < mid.max() < 0.11
    assert np.abs(vy[0, 2:-2, nx // 2]).max() < 1e-3
    assert rho[0, ny // 2, 2] > rho[0, ny // 2, -3] > 0.99          # a pressure drop drives the flow
