"""ctypes binding of libflownet.so (include/flownet.h).  PyTorch is used only for device memory
and streams; every kernel that runs is ours.  There is NO fallback: if the library is missing or
a call fails this module raises."""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FLOWNET_LIB", os.path.join(_HERE, "libflownet.so"))   # override: A/B builds

FN_ABI_VERSION = 10
# enum fn_prologue
PRO_NONE, PRO_CONCAT_ELU, PRO_ELU, PRO_CONCAT_RELU, PRO_RELU = 0, 1, 2, 3, 4
# enum fn_epilogue
EPI_BIAS, EPI_GATE_RES, EPI_RES, EPI_TANH, EPI_GATE_BWD, EPI_PRO_BWD = 0, 1, 2, 3, 4, 5
# device int64 [1] added to every dropout seed inside the kernels (None: none) -- set by flow_net.CompiledTraining so that
# replays of the captured step draw a new mask
DROPOUT_SEED_DEV = None
# enum fn_impl
IMPL_AUTO, IMPL_SIMT, IMPL_TCGEN05 = 0, 1, 2


class ConvDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("Cin", C.c_int32), ("Cout", C.c_int32), ("ksize", C.c_int32), ("stride", C.c_int32),
        ("prologue", C.c_int32), ("epilogue", C.c_int32), ("impl", C.c_int32),
        ("x", C.c_void_p), ("w", C.c_void_p), ("w_tc", C.c_void_p), ("bias", C.c_void_p), ("y", C.c_void_p),
        ("skip", C.c_void_p), ("w_skip", C.c_void_p),
        ("Cskip", C.c_int32), ("SH", C.c_int32), ("SW", C.c_int32),
        ("res", C.c_void_p),
        ("Cres", C.c_int32), ("RH", C.c_int32), ("RW", C.c_int32), ("res_pool", C.c_int32),
        ("keep_prob", C.c_float),
        ("dropout_seed", C.c_uint64), ("dropout_offset", C.c_uint64),
        ("dropout_seed_dev", C.c_void_p),
        ("adj_prologue", C.c_int32), ("adj_accumulate", C.c_int32),
        ("x_planes", C.c_int32), ("skip_planes", C.c_int32), ("y_planes", C.c_void_p),
        ("x_u8", C.c_int32), ("res_u8", C.c_int32),
    ]


class LbmDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("nx", C.c_int32), ("ny", C.c_int32), ("dtype", C.c_int32),
        ("tau", C.c_double), ("u_max", C.c_double), ("rho_out", C.c_double),
        ("solid", C.c_void_p), ("f_a", C.c_void_p), ("f_b", C.c_void_p),
    ]


class Level5Desc(C.Structure):
    _fields_ = [("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32),
                ("x", C.c_void_p), ("w_stream", C.c_void_p), ("bias", C.c_void_p), ("y", C.c_void_p),
                ("y_planes", C.c_void_p), ("split", C.c_int32), ("reserved", C.c_int32)]


class Level4Desc(C.Structure):
    _fields_ = [("B", C.c_int32), ("up", C.c_int32), ("x", C.c_void_p), ("skip", C.c_void_p), ("w_stream", C.c_void_p),
                ("bias", C.c_void_p), ("y", C.c_void_p)]


class Level45Desc(C.Structure):
    _fields_ = [("B", C.c_int32), ("reserved", C.c_int32), ("x", C.c_void_p), ("x4", C.c_void_p), ("z", C.c_void_p),
                ("w_down", C.c_void_p), ("bias_down", C.c_void_p), ("w_level5", C.c_void_p), ("bias_level5", C.c_void_p),
                ("w_up", C.c_void_p), ("bias_up", C.c_void_p), ("y", C.c_void_p)]


class Level3Desc(C.Structure):
    _fields_ = [("B", C.c_int32), ("up", C.c_int32), ("x", C.c_void_p), ("aux", C.c_void_p), ("w_stream", C.c_void_p),
                ("bias", C.c_void_p), ("y", C.c_void_p)]


class TConvDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Cin", C.c_int32), ("Cout", C.c_int32),
        ("impl", C.c_int32),
        ("x", C.c_void_p), ("w", C.c_void_p), ("w_tc", C.c_void_p), ("bias", C.c_void_p), ("y", C.c_void_p),
        ("epilogue", C.c_int32), ("adj_prologue", C.c_int32), ("adj_accumulate", C.c_int32), ("reserved", C.c_int32),
        ("res", C.c_void_p), ("y_planes", C.c_void_p),
    ]


# every symbol include/flownet.h declares: (name, restype, argtypes)
_P = C.c_void_p
SYMBOLS = [
    ("fn_abi_version", C.c_int, []),
    ("fn_sizeof_conv_desc", C.c_int, []),
    ("fn_sizeof_tconv_desc", C.c_int, []),
    ("fn_last_error", C.c_char_p, []),
    ("fn_launch_count", C.c_uint64, []),
    ("fn_last_used_tcgen05", C.c_int, []),
    ("fn_conv_fwd", C.c_int, [C.POINTER(ConvDesc), _P]),
    ("fn_conv_fwd_has_tcgen05", C.c_int, [C.POINTER(ConvDesc)]),
    ("fn_conv_weights_layout", C.c_int, [C.POINTER(ConvDesc)]),
    ("fn_pack_conv_weights_tc_f32", C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    ("fn_tconv_fwd", C.c_int, [C.POINTER(TConvDesc), _P]),
    ("fn_tconv_fwd_has_tcgen05", C.c_int, [C.POINTER(TConvDesc)]),
    ("fn_conv_weights_tc_bytes", C.c_int64, [C.c_int32] * 5),
    ("fn_pack_conv_weights_tc", C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    ("fn_level5_weight_bytes", C.c_int64, []),
    ("fn_level5_split_weight_bytes", C.c_int64, []),
    ("fn_level5_supported", C.c_int, [C.c_int32, C.c_int32, C.c_int32]),
    ("fn_level5_fwd", C.c_int, [C.POINTER(Level5Desc), _P]),
    ("fn_level4_weight_bytes", C.c_int64, [C.c_int32]),
    ("fn_level4_fwd", C.c_int, [C.POINTER(Level4Desc), _P]),
    ("fn_level45_fwd", C.c_int, [C.POINTER(Level45Desc), _P]),
    ("fn_level3_weight_bytes", C.c_int64, [C.c_int32]),
    ("fn_level3_fwd", C.c_int, [C.POINTER(Level3Desc), _P]),
    ("fn_nonlinearity", C.c_int, [_P, _P, C.c_int64, C.c_int32, C.c_int32, _P]),
    ("fn_cast_f32_to_bf16", C.c_int, [_P, _P, C.c_int64, _P]),
    ("fn_cast_u8_to_bf16", C.c_int, [_P, _P, C.c_int64, _P]),
    ("fn_cast_bf16_to_f32", C.c_int, [_P, _P, C.c_int64, _P]),
    ("fn_l2_loss", C.c_int, [_P, _P, _P, _P, C.c_int64, _P]),
    ("fn_adam_step", C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double, _P]),
    ("fn_gate_bwd", C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, _P]),
    ("fn_prologue_bwd_dev", C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_uint64, _P, _P]),
    ("fn_prologue_bwd", C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_uint64, _P]),
    ("fn_nin_bwd", C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    ("fn_shortcut_bwd", C.c_int, [_P, _P] + [C.c_int32] * 9 + [_P]),
    ("fn_tanh_l2_bwd", C.c_int, [_P, _P, _P, _P, C.c_int64, _P]),
    ("fn_bias_grad_workspace", C.c_int64, [C.c_int64, C.c_int32]),
    ("fn_bias_grad", C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, _P]),
    ("fn_conv_wgrad_workspace", C.c_int64, [C.POINTER(ConvDesc)]),
    ("fn_conv_wgrad", C.c_int, [C.POINTER(ConvDesc), _P, _P, _P, _P]),
    ("fn_conv_wgrad_bias", C.c_int, [C.POINTER(ConvDesc), _P, _P, _P, _P, _P]),
    ("fn_tconv_wgrad_workspace", C.c_int64, [C.POINTER(TConvDesc)]),
    ("fn_tconv_wgrad", C.c_int, [C.POINTER(TConvDesc), _P, _P, _P, _P]),
    ("fn_tfrecord_index", C.c_int64, [_P, C.c_int64, _P, _P, C.c_int64, C.c_int32]),
    ("fn_tfrecord_parse_flow_example", C.c_int, [_P, C.c_int64, _P, C.c_int64, _P, C.c_int64]),
    ("fn_tfrecord_parse_flow_batch", C.c_int, [_P, _P, _P, C.c_int64, _P, C.c_int64, _P, C.c_int64, C.c_int32]),
    ("fn_tfrecord_write_flow_example", C.c_int64, [_P, C.c_int64, _P, C.c_int64, _P, C.c_int64]),
    ("fn_lbm_init", C.c_int, [C.POINTER(LbmDesc), C.c_double, C.c_double, C.c_double, _P]),
    ("fn_lbm_steps", C.c_int, [C.POINTER(LbmDesc), C.c_int32, _P]),
    ("fn_lbm_fields", C.c_int, [C.POINTER(LbmDesc), _P, _P, _P]),
    ("fn_crc32c", C.c_uint32, [_P, C.c_int64]),
    ("fn_crc32c_portable", C.c_uint32, [_P, C.c_int64]),
]
DEBUG_SYMBOLS = [
    ("fn_debug_set_swap_lbo_sbo", None, [C.c_int]),
    ("fn_debug_tc_status", C.c_int, []),
    ("fn_debug_set_timing", None, [C.c_void_p, C.c_int]),
    ("fn_debug_set_use_s1p", None, [C.c_int]),
    ("fn_debug_set_wgrad_swap", None, [C.c_int]),
    ("fn_debug_set_nsplit", None, [C.c_int]),
    ("fn_debug_set_wg", None, [C.c_int]),
    ("fn_debug_set_wg_j16", None, [C.c_int]),
    ("fn_debug_set_pl", None, [C.c_int]),
    ("fn_debug_set_pl_grid_cap", None, [C.c_int]),
    ("fn_debug_set_pl_ne", None, [C.c_int]),
    ("fn_debug_set_pl_j", None, [C.c_int]),
    ("fn_debug_set_tail_v2", None, [C.c_int]),
    ("fn_debug_set_pl_ni", None, [C.c_int]),
    ("fn_debug_set_pl_knock", None, [C.c_int]),
]


def _load():
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
            "flow_net has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS + DEBUG_SYMBOLS:
        f = getattr(lib, name)          # AttributeError if the symbol is not exported
        f.restype = res
        f.argtypes = args
    v = lib.fn_abi_version()
    if v != FN_ABI_VERSION:
        raise RuntimeError(f"libflownet.so ABI {v}, binding expects {FN_ABI_VERSION}")
    if lib.fn_sizeof_conv_desc() != C.sizeof(ConvDesc) or lib.fn_sizeof_tconv_desc() != C.sizeof(TConvDesc):
        raise RuntimeError("descriptor struct layout differs between include/flownet.h and the ctypes binding")
    return lib


lib = _load()


class FlownetError(RuntimeError):
    pass


def _check(rc, what):
    if rc == 0:
        return
    msg = lib.fn_last_error().decode("utf-8", "replace")
    if rc < 0:
        raise ValueError(f"{what}: {msg} (code {rc})")
    raise FlownetError(f"{what}: CUDA error {rc}: {msg}")


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise ValueError("flow_net kernels take CUDA tensors; there is no CPU path")


def launch_count() -> int:
    return int(lib.fn_launch_count())


def last_used_tcgen05() -> bool:
    return bool(lib.fn_last_used_tcgen05())


def tc_status() -> int:
    """Synchronises; non-zero means a tcgen05 pipeline wait timed out (debug aid)."""
    return int(lib.fn_debug_tc_status())


def conv_fwd(x, w, bias, y, *, B, H, W, Cin, Cout, ksize=3, stride=1, prologue=PRO_NONE, epilogue=EPI_BIAS,
             impl=IMPL_AUTO, w_tc=None, skip=None, w_skip=None, res=None, res_pool=False, keep_prob=1.0,
             dropout_seed=0, dropout_offset=0, adj_prologue=0, adj_accumulate=False, query=False, x_planes=False,
             skip_planes=False, y_planes=None):
    # DROPOUT_SEED_DEV (module level): optional device int64 [1] added to every dropout seed inside the kernels
    """query=True: do not launch, return whether this call would run on the tensor cores."""
    # x_planes / skip_planes: x / skip are planes tensors [B, 2C/8, H, W, 8] (include/flownet.h); y_planes: optional planes
    # output, y may then be None
    _require_cuda(x, w, bias, y, skip, w_skip, res, w_tc, y_planes)
    d = ConvDesc()
    d.B, d.H, d.W, d.Cin, d.Cout, d.ksize, d.stride = B, H, W, Cin, Cout, ksize, stride
    d.prologue, d.epilogue, d.impl = prologue, epilogue, impl
    d.x, d.w, d.w_tc, d.bias, d.y = x.data_ptr(), w.data_ptr(), (w_tc.data_ptr() if w_tc is not None else None), \
        bias.data_ptr(), (y.data_ptr() if y is not None else None)
    d.x_planes, d.skip_planes = int(bool(x_planes)), int(bool(skip_planes))
    d.y_planes = y_planes.data_ptr() if y_planes is not None else None
    # the boundary image as it arrives (uint8): read directly by the head conv / the first block's one-channel shortcut
    d.x_u8 = int(x.dtype == torch.uint8)
    d.res_u8 = int(res is not None and res.dtype == torch.uint8)
    if skip is not None:
        d.skip, d.w_skip = skip.data_ptr(), w_skip.data_ptr()
        if skip_planes:
            d.Cskip, d.SH, d.SW = skip.shape[1] * 4, skip.shape[2], skip.shape[3]
        else:
            d.Cskip, d.SH, d.SW = skip.shape[3], skip.shape[1], skip.shape[2]
    if res is not None:
        d.res = res.data_ptr()
        d.Cres, d.RH, d.RW = res.shape[3], res.shape[1], res.shape[2]
        d.res_pool = 1 if res_pool else 0
    d.keep_prob = float(keep_prob)
    d.dropout_seed, d.dropout_offset = int(dropout_seed), int(dropout_offset)
    d.dropout_seed_dev = DROPOUT_SEED_DEV.data_ptr() if DROPOUT_SEED_DEV is not None else None
    d.adj_prologue, d.adj_accumulate = int(adj_prologue), int(bool(adj_accumulate))
    if query == "layout":            # 1: the kernel that will run reads the pass-major weight image (fn_conv_weights_layout)
        return int(lib.fn_conv_weights_layout(C.byref(d)))
    if query:
        return bool(lib.fn_conv_fwd_has_tcgen05(C.byref(d)))
    _check(lib.fn_conv_fwd(C.byref(d), stream_ptr()), "fn_conv_fwd")


def tconv_fwd(x, w, bias, y, *, B, H, W, Cin, Cout, impl=IMPL_AUTO, w_tc=None, epilogue=EPI_BIAS, res=None,
              adj_prologue=PRO_NONE, adj_accumulate=False, query=False, y_planes=None):
    """query=True: return whether a tensor-core kernel takes this call (for EPI_PRO_BWD: the fused one), run nothing."""
    _require_cuda(x, w, bias, y, w_tc, res)
    d = TConvDesc()
    d.B, d.H, d.W, d.Cin, d.Cout, d.impl = B, H, W, Cin, Cout, impl
    d.x, d.w, d.bias, d.y = x.data_ptr(), w.data_ptr(), bias.data_ptr(), y.data_ptr()
    d.w_tc = w_tc.data_ptr() if w_tc is not None else None
    d.epilogue, d.adj_prologue, d.adj_accumulate = epilogue, adj_prologue, int(bool(adj_accumulate))
    d.res = res.data_ptr() if res is not None else None
    d.y_planes = y_planes.data_ptr() if y_planes is not None else None
    if query:
        return bool(lib.fn_tconv_fwd_has_tcgen05(C.byref(d)))
    _check(lib.fn_tconv_fwd(C.byref(d), stream_ptr()), "fn_tconv_fwd")


def level5_supported(H, W, Cc):
    return bool(lib.fn_level5_supported(int(H), int(W), int(Cc)))


def level5_fwd(x, w_stream, bias, y, y_planes=None, split=False):
    """x: bf16 [B,16,32,64] -> y: bf16 [B,16,32,64]: the 8x16 level in one launch (fn_level5_fwd).  split: two CTAs per
    image, w_stream / bias in the per-rank form (include/flownet.h)"""
    _require_cuda(x, w_stream, bias, y, y_planes)
    want = int(lib.fn_level5_split_weight_bytes() if split else lib.fn_level5_weight_bytes())
    if w_stream.numel() * w_stream.element_size() != want or bias.numel() != 832:
        raise ValueError("level5_fwd: weight stream / bias size")
    d = Level5Desc()
    d.B, d.H, d.W, d.C = x.shape
    d.x, d.w_stream, d.bias, d.y = x.data_ptr(), w_stream.data_ptr(), bias.data_ptr(), y.data_ptr()
    d.y_planes = y_planes.data_ptr() if y_planes is not None else None
    d.split = int(bool(split))
    _check(lib.fn_level5_fwd(C.byref(d), stream_ptr()), "fn_level5_fwd")


def level4_fwd(x, w_stream, bias, y, skip=None):
    """The 16x32 level around fn_level5_fwd (fn_level4_fwd, csrc/conv_l4.cu).  skip is None: x = x3 bf16 [B,32,64,32] ->
    y = x4 bf16 [B,16,32,64] (resnet_4_downsample + resnet_4_0); with skip = x4: x = up_conv_1's output [B,16,32,64] ->
    y bf16 [B,32,64,32] (resnet_up_1_0 + up_conv_2)."""
    _require_cuda(x, w_stream, bias, y, skip)
    up = skip is not None
    want_x, want_y = ((16, 32, 64), (32, 64, 32)) if up else ((32, 64, 32), (16, 32, 64))
    if tuple(x.shape[1:]) != want_x or tuple(y.shape[1:]) != want_y or y.shape[0] != x.shape[0] or \
            (up and tuple(skip.shape) != tuple(x.shape)):
        raise ValueError("level4_fwd: tensor shapes")
    if w_stream.numel() * w_stream.element_size() != int(lib.fn_level4_weight_bytes(int(up))) or bias.numel() != (224 if up else 384):
        raise ValueError("level4_fwd: weight stream / bias size")
    d = Level4Desc()
    d.B, d.up = x.shape[0], int(up)
    d.x, d.w_stream, d.bias, d.y = x.data_ptr(), w_stream.data_ptr(), bias.data_ptr(), y.data_ptr()
    d.skip = skip.data_ptr() if up else None
    _check(lib.fn_level4_fwd(C.byref(d), stream_ptr()), "fn_level4_fwd")


def level45_fwd(x3, x4, z, w_down, bias_down, w5, bias5, w_up, bias_up, y):
    """Levels 4 and 5 in one launch (fn_level45_fwd): x3 bf16 [B,32,64,32] -> y bf16 [B,32,64,32] (output of up_conv_2); x4 / z:
    bf16 [B,16,32,64] scratch; the weight streams / biases of level4_fwd (down), level5_fwd (split form) and level4_fwd (up)."""
    _require_cuda(x3, x4, z, w_down, bias_down, w5, bias5, w_up, bias_up, y)
    B = x3.shape[0]
    if tuple(x3.shape) != (B, 32, 64, 32) or tuple(y.shape) != (B, 32, 64, 32) or tuple(x4.shape) != (B, 16, 32, 64) or tuple(z.shape) != (B, 16, 32, 64):
        raise ValueError("level45_fwd: tensor shapes")
    nb = lambda t: t.numel() * t.element_size()
    if nb(w_down) != int(lib.fn_level4_weight_bytes(0)) or nb(w_up) != int(lib.fn_level4_weight_bytes(1)) or \
            nb(w5) != int(lib.fn_level5_split_weight_bytes()) or bias_down.numel() != 384 or bias_up.numel() != 224 or bias5.numel() != 832:
        raise ValueError("level45_fwd: weight stream / bias size")
    d = Level45Desc()
    d.B = B
    d.x, d.x4, d.z, d.y = x3.data_ptr(), x4.data_ptr(), z.data_ptr(), y.data_ptr()
    d.w_down, d.bias_down = w_down.data_ptr(), bias_down.data_ptr()
    d.w_level5, d.bias_level5 = w5.data_ptr(), bias5.data_ptr()
    d.w_up, d.bias_up = w_up.data_ptr(), bias_up.data_ptr()
    _check(lib.fn_level45_fwd(C.byref(d), stream_ptr()), "fn_level45_fwd")


def level3_fwd(x1, aux, w_stream, bias, y, up):
    """The stride-1 convolutions of the 32x64 level (fn_level3_fwd, csrc/conv_l3.cu).  up False: x1 = conv_1 output of
    resnet_3_downsample [B,32,64,32], aux = x2 [B,64,128,16] -> y = x3 [B,32,64,32]; up True: x1 = conv_1 (+ nin) output of
    resnet_up_2_0, aux = the block's input z [B,32,64,32] -> y [B,64,128,16] (output of up_conv_3)."""
    _require_cuda(x1, aux, w_stream, bias, y)
    want_aux, want_y = ((32, 64, 32), (64, 128, 16)) if up else ((64, 128, 16), (32, 64, 32))
    if tuple(x1.shape[1:]) != (32, 64, 32) or tuple(aux.shape[1:]) != want_aux or tuple(y.shape[1:]) != want_y or \
            not (x1.shape[0] == aux.shape[0] == y.shape[0]):
        raise ValueError("level3_fwd: tensor shapes")
    if w_stream.numel() * w_stream.element_size() != int(lib.fn_level3_weight_bytes(int(up))) or bias.numel() != (80 if up else 160):
        raise ValueError("level3_fwd: weight stream / bias size")
    d = Level3Desc()
    d.B, d.up = x1.shape[0], int(up)
    d.x, d.aux, d.w_stream, d.bias, d.y = x1.data_ptr(), aux.data_ptr(), w_stream.data_ptr(), bias.data_ptr(), y.data_ptr()
    _check(lib.fn_level3_fwd(C.byref(d), stream_ptr()), "fn_level3_fwd")


def pack_conv_weights_tc(w, w_skip, ksize, Keff, Kskip, Cout, gated):
    """w: bf16 [k*k, Keff, Cout] (+ w_skip bf16 [Kskip, Cout]) -> tensor-core image (bf16, 1-D) or None."""
    nbytes = int(lib.fn_conv_weights_tc_bytes(ksize, Keff, Kskip, Cout, int(gated)))
    if nbytes == 0:
        return None
    dst = torch.empty(nbytes // 2, dtype=torch.bfloat16, device=w.device)
    _check(lib.fn_pack_conv_weights_tc(_ptr(w), _ptr(w_skip), _ptr(dst), ksize, Keff, Kskip, Cout, int(gated),
                                       stream_ptr()), "fn_pack_conv_weights_tc")
    return dst


def pack_conv_weights_tc_f32(w, ksize, K, Cout, flip_taps=False, transpose=False):
    """tensor-core weight image from an fp32 variable [k*k (or k,k), A, B]; see fn_pack_conv_weights_tc_f32"""
    _require_cuda(w)
    nbytes = lib.fn_conv_weights_tc_bytes(ksize, K, 0, Cout, 0)
    if nbytes <= 0:
        raise ValueError(f"no tensor-core weight image for K={K}")
    dst = torch.empty(nbytes // 2, dtype=torch.bfloat16, device=w.device)
    _check(lib.fn_pack_conv_weights_tc_f32(_ptr(w), _ptr(dst), ksize, K, Cout, int(flip_taps), int(transpose), stream_ptr()),
           "fn_pack_conv_weights_tc_f32")
    return dst


def nonlinearity(x, kind):
    _require_cuda(x)
    Cc = x.shape[-1]
    mult = 2 if kind in (PRO_CONCAT_ELU, PRO_CONCAT_RELU) else 1
    y = torch.empty(tuple(x.shape[:-1]) + (mult * Cc,), dtype=torch.bfloat16, device=x.device)
    n = x.numel() // Cc
    _check(lib.fn_nonlinearity(_ptr(x), _ptr(y), n, Cc, kind, stream_ptr()), "fn_nonlinearity")
    return y


def to_bf16(x):
    """Boundary plumbing: uint8 / float32 / bf16 CUDA tensor -> contiguous bf16."""
    _require_cuda(x)
    x = x.contiguous()
    if x.dtype == torch.bfloat16:
        return x
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    if x.dtype == torch.float32:
        _check(lib.fn_cast_f32_to_bf16(_ptr(x), _ptr(y), x.numel(), stream_ptr()), "fn_cast_f32_to_bf16")
    elif x.dtype == torch.uint8:
        _check(lib.fn_cast_u8_to_bf16(_ptr(x), _ptr(y), x.numel(), stream_ptr()), "fn_cast_u8_to_bf16")
    else:
        raise ValueError(f"unsupported boundary dtype {x.dtype} (use uint8, float32 or bfloat16)")
    return y


def to_f32(x):
    _require_cuda(x)
    x = x.contiguous()
    if x.dtype == torch.float32:
        return x
    if x.dtype != torch.bfloat16:
        raise ValueError(f"unsupported dtype {x.dtype}")
    y = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    _check(lib.fn_cast_bf16_to_f32(_ptr(x), _ptr(y), x.numel(), stream_ptr()), "fn_cast_bf16_to_f32")
    return y


def l2_loss(p, t, want_grad=False):
    """tf.nn.l2_loss(p - t): returns (loss[1] fp32, grad or None)."""
    _require_cuda(p, t)
    p = p.contiguous()
    t = t.contiguous()
    if p.dtype != torch.float32 or t.dtype != torch.float32 or p.shape != t.shape:
        raise ValueError("l2_loss takes two float32 tensors of the same shape")
    loss = torch.zeros(1, dtype=torch.float32, device=p.device)
    grad = torch.empty_like(p) if want_grad else None
    _check(lib.fn_l2_loss(_ptr(p), _ptr(t), _ptr(loss), _ptr(grad), p.numel(), stream_ptr()), "fn_l2_loss")
    return loss, grad


def adam_step(param, grad, m, v, step, lr=1e-4, beta1=0.9, beta2=0.999, eps=1e-8):
    _require_cuda(param, grad, m, v)
    for t in (param, grad, m, v):
        if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != param.numel():
            raise ValueError("adam_step takes flat contiguous float32 buffers of equal size")
    _check(lib.fn_adam_step(_ptr(param), _ptr(grad), _ptr(m), _ptr(v), param.numel(), int(step), lr, beta1, beta2,
                            eps, stream_ptr()), "fn_adam_step")


# ------------------------------------------------------------------ training-step adjoints
_ws_cache = {}
_ws_retired = []     # outgrown buffers are kept: a captured CUDA graph may have their address baked in


def _workspace(nfloats, device):
    """One grow-only fp32 scratch buffer per device for the two-stage reductions.  A buffer that is outgrown is retired,
    never freed: replays of a graph captured earlier (flow_net.CompiledTraining) keep writing through the old pointer."""
    buf = _ws_cache.get(device)
    if buf is None or buf.numel() < nfloats:
        if buf is not None:
            _ws_retired.append(buf)
        buf = torch.empty(max(int(nfloats) * 3 // 2, 1 << 20), dtype=torch.float32, device=device)
        _ws_cache[device] = buf
    return buf


def gate_bwd(c2, g):
    _require_cuda(c2, g)
    F = g.shape[-1]
    dc2 = torch.empty_like(c2)
    _check(lib.fn_gate_bwd(_ptr(c2), _ptr(g), _ptr(dc2), g.numel() // F, F, stream_ptr()), "fn_gate_bwd")
    return dc2


def prologue_bwd(dA, x, kind, out=None, accumulate=False, keep_prob=1.0, seed=0):
    _require_cuda(dA, x, out)
    Cc = x.shape[-1]
    if out is None:
        out = torch.empty_like(x)
        accumulate = False
    _check(lib.fn_prologue_bwd_dev(_ptr(dA), _ptr(x), _ptr(out), x.numel() // Cc, Cc, kind, int(accumulate), float(keep_prob),
                                   int(seed), _ptr(DROPOUT_SEED_DEV), stream_ptr()), "fn_prologue_bwd")
    return out


def nin_bwd(dy, w, a, kind, out=None, accumulate=False):
    """d(a) (+)= adjoint of pro(a) applied to dy @ w^T; w: the fp32 [2*Ca, F] nin variable"""
    _require_cuda(dy, w, a, out)
    if out is None:
        out = torch.empty_like(a)
        accumulate = False
    Ca, F = a.shape[-1], dy.shape[-1]
    _check(lib.fn_nin_bwd(_ptr(dy), _ptr(w), _ptr(a), _ptr(out), a.numel() // Ca, F, Ca, kind, int(accumulate), stream_ptr()),
           "fn_nin_bwd")
    return out


def shortcut_bwd(g, x_shape, pooled, out=None, accumulate=False):
    _require_cuda(g, out)
    B, H, W, Cin = x_shape
    _, OH, OW, F = g.shape
    if out is None:
        out = torch.empty(tuple(x_shape), dtype=torch.bfloat16, device=g.device)
        accumulate = False
    _check(lib.fn_shortcut_bwd(_ptr(g), _ptr(out), B, OH, OW, F, H, W, Cin, int(pooled), int(accumulate), stream_ptr()),
           "fn_shortcut_bwd")
    return out


def tanh_l2_bwd(p, t):
    """returns (loss [1] fp32, d(pre-tanh) bf16 [..., 8] with channels 2..7 zero)"""
    _require_cuda(p, t)
    p = p.contiguous()
    t = t.contiguous()
    npix = p.numel() // 2
    loss = torch.zeros(1, dtype=torch.float32, device=p.device)
    dpre = torch.empty(tuple(p.shape[:-1]) + (8,), dtype=torch.bfloat16, device=p.device)
    _check(lib.fn_tanh_l2_bwd(_ptr(p), _ptr(t), _ptr(loss), _ptr(dpre), npix, stream_ptr()), "fn_tanh_l2_bwd")
    return loss, dpre


def bias_grad(dY, out):
    _require_cuda(dY, out)
    N = dY.shape[-1]
    npix = dY.numel() // N
    ws = _workspace(lib.fn_bias_grad_workspace(npix, N), dY.device)
    _check(lib.fn_bias_grad(_ptr(dY), _ptr(out), _ptr(ws), npix, N, stream_ptr()), "fn_bias_grad")


def conv_wgrad(x, dY, out, *, ksize, stride, prologue, keep_prob=1.0, dropout_seed=0, impl=IMPL_AUTO, db=None):
    """out: fp32 [k*k, Keff, Cout] <- sum over pixels of pro(x) (x) dY; db (fp32 [Cout], optional) <- sum over pixels of dY"""
    _require_cuda(x, dY, out, db)
    d = ConvDesc()
    d.impl = impl
    d.B, d.H, d.W, d.Cin = x.shape
    d.Cout = dY.shape[-1]
    d.ksize, d.stride, d.prologue = ksize, stride, prologue
    d.x = x.data_ptr()
    d.keep_prob = float(keep_prob)
    d.dropout_seed = int(dropout_seed)
    d.dropout_seed_dev = DROPOUT_SEED_DEV.data_ptr() if DROPOUT_SEED_DEV is not None else None
    ws = _workspace(lib.fn_conv_wgrad_workspace(C.byref(d)), x.device)
    if db is not None:
        _check(lib.fn_conv_wgrad_bias(C.byref(d), _ptr(dY), _ptr(out), _ptr(db), _ptr(ws), stream_ptr()), "fn_conv_wgrad_bias")
    else:
        _check(lib.fn_conv_wgrad(C.byref(d), _ptr(dY), _ptr(out), _ptr(ws), stream_ptr()), "fn_conv_wgrad")


def tconv_wgrad(x, dY, out, *, impl=IMPL_AUTO):
    """out: fp32 [9, Cout, Cin] -- the variable's own [k,k,Cout,Cin] layout"""
    _require_cuda(x, dY, out)
    d = TConvDesc()
    d.impl = impl
    d.B, d.H, d.W, d.Cin = x.shape
    d.Cout = dY.shape[-1]
    d.x = x.data_ptr()
    ws = _workspace(lib.fn_tconv_wgrad_workspace(C.byref(d)), x.device)
    _check(lib.fn_tconv_wgrad(C.byref(d), _ptr(dY), _ptr(out), _ptr(ws), stream_ptr()), "fn_tconv_wgrad")
