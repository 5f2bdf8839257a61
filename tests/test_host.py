"""CPU-side tests of the host logic: the C-ABI library builds, loads and exports every symbol
include/flownet.h declares (no compute calls -- there is no GPU here), the ctypes structs match the
C layout, argument validation answers without touching the device, and the product's variable
store / synthetic inputs follow the same conventions as the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from oracle import flow_oracle as fo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, "include", "flownet.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(fn_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 15
    bound = {n for n, _, _ in built_lib.SYMBOLS}
    assert declared == bound, (declared ^ bound)
    raw = C.CDLL(built_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), name


def test_abi_version_and_struct_sizes(built_lib):
    assert built_lib.lib.fn_abi_version() == built_lib.FN_ABI_VERSION
    assert built_lib.lib.fn_sizeof_conv_desc() == C.sizeof(built_lib.ConvDesc)
    assert built_lib.lib.fn_sizeof_tconv_desc() == C.sizeof(built_lib.TConvDesc)


def test_validation_errors_without_device(built_lib):
    lib = built_lib.lib
    assert lib.fn_conv_fwd(None, None) == -1
    assert b"null descriptor" in lib.fn_last_error()
    d = built_lib.ConvDesc()
    assert lib.fn_conv_fwd(C.byref(d), None) == -1          # non-positive dims
    d.B, d.H, d.W, d.Cin, d.Cout, d.ksize, d.stride = 1, 8, 8, 12, 8, 3, 1
    d.x = d.w = d.bias = d.y = 4096
    assert lib.fn_conv_fwd(C.byref(d), None) == -3          # Cin not 1 / multiple of 8
    d.Cin, d.ksize = 8, 5
    assert lib.fn_conv_fwd(C.byref(d), None) == -2          # unsupported kernel size
    d.ksize, d.x = 3, 4097
    assert lib.fn_conv_fwd(C.byref(d), None) == -3          # misaligned pointer
    assert lib.fn_tconv_fwd(None, None) == -1
    t = built_lib.TConvDesc()
    t.B, t.H, t.W, t.Cin, t.Cout = 1, 16, 16, 16, 8
    t.x = t.w = t.bias = t.y = 4096
    t.epilogue = 7
    assert lib.fn_tconv_fwd(C.byref(t), None) == -1         # unknown epilogue
    t.epilogue, t.adj_prologue = built_lib.EPI_PRO_BWD, built_lib.PRO_CONCAT_ELU
    assert lib.fn_tconv_fwd(C.byref(t), None) == -2         # prologue-adjoint epilogue without res / with Cout != Cin
    assert lib.fn_tconv_fwd_has_tcgen05(C.byref(t)) == 0
    assert lib.fn_tconv_fwd_has_tcgen05(None) == 0
    assert lib.fn_adam_step(None, None, None, None, 1, 1, 1e-4, 0.9, 0.999, 1e-8, None) == -1
    assert lib.fn_l2_loss(None, None, None, None, 4, None) == -1


def test_cpu_tensors_are_rejected_not_silently_computed(built_lib):
    with pytest.raises(ValueError):
        built_lib.to_bf16(torch.zeros(4, 8))
    with pytest.raises(ValueError):
        built_lib.l2_loss(torch.zeros(4), torch.zeros(4))


def test_tc_weight_image_size(built_lib):
    lib = built_lib.lib
    # r1_0.conv_2: 3x3, Keff=16, N=16 -> 9 k-steps * 16 * 32 B
    assert lib.fn_conv_weights_tc_bytes(3, 16, 0, 16, 1) == 9 * 16 * 32
    # ru4_0.conv_1 + nin: Cout=8 padded to N=16, 9 + 1 k-steps
    assert lib.fn_conv_weights_tc_bytes(3, 16, 16, 8, 0) == 10 * 16 * 32
    # K = 8 (8-channel gradients): one k-step per tap whose second half is zero
    assert lib.fn_conv_weights_tc_bytes(3, 8, 0, 8, 0) == 9 * 16 * 32
    assert lib.fn_conv_weights_tc_bytes(3, 24, 0, 8, 0) == 0    # K not a multiple of 16: no tcgen05 image


def test_product_variable_store_matches_oracle_convention(built_lib):
    import model.flow_architecture as arch
    vs = fo.VariableStore(1234, torch.float32)
    fo.inference(vs, torch.zeros(1, 16, 32, 1), 1.0)
    arch.reset_variables(1234)
    for name, v in vs.vars.items():
        pv = arch._variable(name, list(v.shape), device="cpu")
        assert torch.equal(pv, v), name
    assert list(arch.global_variables()) == list(vs.vars)
    arch.reset_variables(1234)


def test_product_synthetic_inputs_match_oracle(built_lib):
    from utils.synthetic import synth_boundary
    for (B, H, W, seed) in [(2, 128, 256, 0), (1, 16, 32, 0), (2, 32, 48, 5)]:
        np.testing.assert_array_equal(synth_boundary(B, H, W, seed)[..., 0],
                                      fo.synth_boundary(B, H, W, seed)[..., 0].astype(np.uint8))


def test_flags_defaults_match_reference(built_lib):
    import model.flow_net as flow_net
    F = flow_net.FLAGS
    assert (F.model, F.nr_res_blocks, F.gated_res, F.nonlinearity) == ("res", 1, True, "concat_elu")
    import model.flow_architecture as arch
    with pytest.raises(ValueError):
        arch.set_nonlinearity("swish")
    assert arch.set_nonlinearity("concat_elu") is arch.concat_elu


# ---------------------------------------------------------------- checkpoints keyed by TF names
def test_checkpoint_roundtrip_and_lazy_restore(tmp_path):
    import numpy as np
    import torch
    import model.flow_architecture as arch
    from utils import checkpoint

    class Flat:                      # the fields of flow_backward.FlatParams that a checkpoint touches
        pass

    arch.reset_variables(7)
    a = arch._variable('resnet_1_0_conv_1_conv/weights', [3, 3, 1, 8], device="cpu")
    b = arch._variable('resnet_1_0_conv_1_conv/biases', [8], device="cpu")
    flat = Flat()
    flat.offsets = {'resnet_1_0_conv_1_conv/weights': (0, 72, (3, 3, 1, 8)), 'resnet_1_0_conv_1_conv/biases': (72, 8, (8,))}
    flat.m = torch.arange(80, dtype=torch.float32)
    flat.v = torch.arange(80, dtype=torch.float32) * 2
    flat.step = 41
    path = checkpoint.save(str(tmp_path), 1000, arch.global_variables(), flat)
    assert checkpoint.get_checkpoint_state(str(tmp_path)) == path
    ck = checkpoint.load(path)
    assert set(ck) == {'resnet_1_0_conv_1_conv/weights', 'resnet_1_0_conv_1_conv/biases',
                       'resnet_1_0_conv_1_conv/weights/Adam', 'resnet_1_0_conv_1_conv/weights/Adam_1',
                       'resnet_1_0_conv_1_conv/biases/Adam', 'resnet_1_0_conv_1_conv/biases/Adam_1',
                       'beta1_power', 'beta2_power', 'adam_step', 'global_step'}
    assert ck['resnet_1_0_conv_1_conv/weights'].shape == (3, 3, 1, 8)          # TF HWIO layout, untouched
    wa, wb = a.clone(), b.clone()

    # restore into a fresh graph: the first variable exists already, the second is created lazily afterwards
    arch.reset_variables(99)
    a2 = arch._variable('resnet_1_0_conv_1_conv/weights', [3, 3, 1, 8], device="cpu")
    assert not torch.equal(a2, wa)
    flat2 = Flat()
    flat2.offsets, flat2.m, flat2.v, flat2.step = flat.offsets, torch.zeros(80), torch.zeros(80), 0
    step = checkpoint.restore(path, arch.VARIABLES, flat2, strict=False)
    b2 = arch._variable('resnet_1_0_conv_1_conv/biases', [8], device="cpu")
    assert step == 1000 and flat2.step == 41
    assert torch.equal(a2, wa) and torch.equal(b2, wb)
    assert torch.equal(flat2.m, flat.m) and torch.equal(flat2.v, flat.v)

    # a checkpoint that does not fit the graph is refused (the reference then re-initialises, flow_train.py:66-72)
    arch.reset_variables(5)
    arch._variable('resnet_1_0_conv_1_conv/weights', [3, 3, 1, 16], device="cpu")
    with pytest.raises(ValueError):
        checkpoint.restore(path, arch.VARIABLES)
    arch.reset_variables(1234)


def test_experiment_manager_paths_follow_the_flags(tmp_path):
    from model import flags
    from utils import experiment_manager as em
    import model.flow_net  # noqa: F401  (registers the model flags)
    p = em.make_checkpoint_path(str(tmp_path), flags.FLAGS)
    assert 'nr_res_blocks_1' in p and 'nonlinearity_concat_elu' in p and 'base_dir' not in p
    import os
    os.makedirs(p)
    open(os.path.join(p, 'checkpoint'), 'w').write('model_checkpoint_path: "model.ckpt-0"\n')
    assert em.list_all_checkpoints(str(tmp_path)) == [p]
    old = flags.FLAGS.nr_res_blocks
    try:
        em.set_flags_given_checkpoint_path(p.replace('nr_res_blocks_1', 'nr_res_blocks_3')[len(str(tmp_path)):], flags.FLAGS)
        assert flags.FLAGS.nr_res_blocks == 3
    finally:
        flags.FLAGS.nr_res_blocks = old


def test_weight_prep_is_transparent_outside_a_capture(built_lib):
    """flow_architecture.weight_prep / flow_backward._pad_channels (host logic of the training graphs' side streams): with no
    capture in progress the image is made in place, on the caller's stream, and returned unchanged."""
    import model.flow_architecture as arch
    import model.flow_backward as fb
    assert arch._PREP["stream"] is None
    made = []
    out = arch.weight_prep(lambda: made.append(1) or "image")
    assert out == "image" and made == [1]
    w = torch.arange(3 * 3 * 4 * 2, dtype=torch.float32).reshape(3, 3, 4, 2)
    w8 = fb._pad_channels(w, 3, 8)
    assert w8.shape == (3, 3, 4, 8) and torch.equal(w8[..., :2], w) and float(w8[..., 2:].abs().sum()) == 0.0
    h8 = fb._pad_channels(torch.ones(3, 3, 1, 8), 2, 8)
    assert h8.shape == (3, 3, 8, 8) and float(h8.sum()) == 72.0
    # the fused CUDA-core nin adjoint only where it measured faster than the tensor-core route (tools/ab_nin_bwd.py)
    assert set(fb.NIN_BWD_FUSED_WIDTHS) <= {8, 16, 32, 64}
