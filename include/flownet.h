/* flownet.h -- C ABI of libflownet.so, the sm_100a implementation of the flow_net hot path.
 *
 * The reference (loliverhennigh/Steady-State-Flow-With-Neural-Nets) has no FFI: its hot path
 * is a set of Python functions that emit TensorFlow ops.  Each entry point below replaces the
 * TF op group named beside it (file:line in /root/reference); the Python host code in
 * steady-state-flow-with-neural-nets_b200/model/ keeps the reference's function names and
 * binds these symbols with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer borrowed from the caller unless the name ends in _host
 *   - activations: bf16, NHWC, densely packed; final velocity field: fp32 NHWC
 *   - every call is stream-ordered on `stream` (a cudaStream_t passed as void*), re-entrant,
 *     allocates nothing and never synchronises
 *   - return value: 0 ok; <0 invalid argument / unsupported shape (FN_E_*); >0 a cudaError_t
 *   - fn_last_error() returns a thread-local message for the last non-zero return
 *   - no C++ exception crosses this boundary; there is no CPU fallback
 */
#ifndef FLOWNET_H_
#define FLOWNET_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FN_ABI_VERSION 10

/* error codes (negative) */
#define FN_E_INVAL      (-1)   /* null pointer, bad enum, non-positive dim */
#define FN_E_UNSUPPORTED (-2)  /* shape/flag combination this build has no kernel for */
#define FN_E_ALIGN      (-3)   /* pointer not 16-byte aligned / channel count not a multiple of 8 */

/* A-operand prologue: the nonlinearity the reference applies to the conv input
 * (flow_architecture.py:14-29 set_nonlinearity; res_block :132-135, :142-143) */
enum fn_prologue {
  FN_PRO_NONE        = 0,  /* raw input (first block with 1 channel :132-133; last_conv :242; up-convs) */
  FN_PRO_CONCAT_ELU  = 1,  /* [elu(x), elu(-x)], C -> 2C  (:14-17) */
  FN_PRO_ELU         = 2,  /* tf.nn.elu   (:22-23) */
  FN_PRO_CONCAT_RELU = 3,  /* tf.nn.crelu (:24-25) */
  FN_PRO_RELU        = 4   /* tf.nn.relu  (:26-27) */
};

/* epilogue fused behind the accumulator */
enum fn_epilogue {
  FN_EPI_BIAS     = 0,  /* y = acc + bias                        -> bf16 [.,Cout]      (conv_layer :64-65) */
  FN_EPI_GATE_RES = 1,  /* y = sc(res) + (u+bu)*sigmoid(v+bv)    -> bf16 [.,Cout/2]    (res_block :149-165) */
  FN_EPI_RES      = 2,  /* y = sc(res) + acc + bias              -> bf16 [.,Cout]      (res_block :147,:153-165) */
  FN_EPI_TANH     = 3,  /* y = tanh(acc + bias)                  -> fp32 [.,Cout]      (conv_res :242-243) */
  /* backward of the gate, fused with the recomputation of its input: with [u|v] = acc + bias and g = res (the
   * gradient of the block output, bf16 [.,Cout/2]):  y = [g*sigmoid(v) | g*u*sigmoid(v)*(1-sigmoid(v))] -> bf16 [.,Cout],
   * the gradient of the pre-gate tensor (adjoint of :150-151).  tcgen05 kernels only (concat_elu prologue, stride 1). */
  FN_EPI_GATE_BWD = 4,
  /* backward of a prologue nonlinearity fused into the data-gradient conv that produces its input gradient: acc = dA,
   * the gradient of pro(z) (Cout = mult * Cz channels); res = z, bf16 [.,Cz];
   *   y[., c] (+)= dA[c] * f'(z[c])  -  dA[Cz + c] * f'(-z[c])        (second term: concat kinds only)
   * with the forward's dropout mask (keep_prob, dropout_seed) applied to dA; kind = adj_prologue, += if adj_accumulate.
   * tcgen05 kernels only (stride 1, 3x3, no prologue on x, Cin in {8,16,32,64}, Cout in {Cin, 2 Cin}). */
  FN_EPI_PRO_BWD = 5
};

/* implementation selector: the tensor-core kernels are the product; the CUDA-core kernels
 * cover the layers with no GEMM shape (1-channel head, 2-channel tail) and serve as an
 * on-device cross-check.  FN_IMPL_AUTO picks tcgen05 wherever a kernel exists. */
enum fn_impl { FN_IMPL_AUTO = 0, FN_IMPL_SIMT = 1, FN_IMPL_TCGEN05 = 2 };

/* One fused convolution launch.
 *
 * Computes, for every output pixel p of a TF-'SAME' k x k convolution with stride s
 * (tf.nn.conv2d, flow_architecture.py:64; padding rule SURVEY App. B.1):
 *
 *   acc[p, :] = sum_taps pro(x)[p*s + tap - pad, :] . W[tap]            (K = Keff per tap)
 *             + pro(skip)[p, :] . Wskip                                 (optional: nin, :136-142)
 *   y[p, :]   = epilogue(acc[p, :] + bias, res)
 *
 * where pro() is the prologue nonlinearity and Keff = Cin (x2 for the concat variants).
 * `skip` (if non-null) has the OUTPUT spatial size... or smaller: it is zero-padded at the
 * bottom/right to OH x OW exactly as tf.pad does at :139-141.
 * `res` is the block's orig_x: [B,RH,RW,Cres]; if res_pool it is 2x2/2 'SAME' average-pooled
 * (:153-155) and if Cres < F it is zero-padded IN FRONT on the channel axis (:158-163).
 */
typedef struct fn_conv_desc {
  int32_t B, H, W;          /* input spatial size */
  int32_t Cin;              /* raw input channels (1, or a multiple of 8) */
  int32_t Cout;             /* accumulator width N (gated: 2F) */
  int32_t ksize;            /* 3 or 1 */
  int32_t stride;           /* 1 or 2 */
  int32_t prologue;         /* enum fn_prologue */
  int32_t epilogue;         /* enum fn_epilogue */
  int32_t impl;             /* enum fn_impl */
  const void*  x;           /* bf16 [B,H,W,Cin] */
  const void*  w;           /* bf16 [k*k, Keff, Cout]  (== TF HWIO flattened), SIMT layout   */
  const void*  w_tc;        /* bf16 tensor-core packing of the same weights (fn_pack_conv_weights_tc) or NULL */
  const float* bias;        /* fp32 [Cout] (conv bias, plus nin bias when skip is used) */
  void*        y;           /* bf16 [B,OH,OW,Cy]; fp32 for FN_EPI_TANH */
  const void*  skip;        /* bf16 [B,SH,SW,Cskip] or NULL */
  const void*  w_skip;      /* bf16 [Kskip_eff, Cout] (== TF [dim,hiddens]) */
  int32_t Cskip, SH, SW;
  const void*  res;         /* bf16 [B,RH,RW,Cres] or NULL */
  int32_t Cres, RH, RW;
  int32_t res_pool;         /* 1: 2x2 average pool of res */
  /* training only (tf.nn.dropout on the prologue output, :144-145); keep_prob >= 1 disables */
  float    keep_prob;
  uint64_t dropout_seed;
  uint64_t dropout_offset;
  /* optional device-resident addend to dropout_seed (NULL: none): a captured CUDA graph of the training step gets a new
   * mask on every replay by bumping this counter instead of changing a launch argument */
  const uint64_t* dropout_seed_dev;
  /* FN_EPI_PRO_BWD only */
  int32_t  adj_prologue;    /* enum fn_prologue of the nonlinearity being differentiated */
  int32_t  adj_accumulate;  /* 1: y += ... */
  /* MMA-ready activations ("planes", inference fast path, csrc/conv_pl.cu).  The planes form of a raw tensor
   * t = bf16 [B,H,W,C] is bf16 [B, 2C/8, H, W, 8] = concat_elu(t) (flow_architecture.py:14-17) in 8-channel groups:
   * plane q < C/8 holds elu(t[..., 8q:8q+8]), plane C/8 + q holds elu(-t[..., 8q:8q+8]).  A producer's epilogue writes it
   * once; the consuming convolution's tensor copies then land the tensor-core A operand in shared memory directly. */
  int32_t  x_planes;        /* 1: x is given in planes form (prologue must be FN_PRO_CONCAT_ELU -- it has been applied);
                                  stride 1, 3x3, Cin in {8,16,32}, keep_prob 1, FN_EPI_BIAS / FN_EPI_GATE_RES only */
  int32_t  skip_planes;     /* 1: skip likewise (required with x_planes) */
  void*    y_planes;        /* optional second output: concat_elu(y) in planes form [B, 2Cy/8, OH, OW, 8]; y may then be
                                  NULL.  Taken by the planes kernel, the stride-2 tensor-core kernels and the 1-channel head */
  /* the network input as it arrives: x (with Cin == 1, FN_PRO_NONE) and / or res (with Cres == 1) are uint8 [B,H,W,1] boundary
   * masks instead of bf16 -- the head conv and the first block's one-channel shortcut then read the image directly and the
   * separate cast launch disappears (inference; kernels without a uint8 path answer FN_E_UNSUPPORTED) */
  int32_t  x_u8;
  int32_t  res_u8;
} fn_conv_desc;

/* One transposed convolution launch: tf.nn.conv2d_transpose(k=3, s=2, 'SAME') + bias
 * (flow_architecture.py:70-87; phase taps SURVEY App. B.2).  x: bf16 [B,h,w,Cin] ->
 * y: bf16 [B,2h,2w,Cout].  w: bf16 [9, Cin, Cout] (TF [kh,kw,Cout,Cin] with the last two
 * axes swapped), w_tc: tensor-core packing or NULL. */
typedef struct fn_tconv_desc {
  int32_t B, H, W, Cin, Cout;
  int32_t impl;
  const void*  x;
  const void*  w;
  const void*  w_tc;
  const float* bias;
  void*        y;
  /* FN_EPI_BIAS, or FN_EPI_PRO_BWD (Cout == Cin only, tensor cores only): this transposed conv is the data gradient of a
   * stride-2 conv whose input went through prologue `adj_prologue` (concat kinds); y: bf16 [B,2h,2w,Cout/2] (+)= the
   * prologue's adjoint at res = the differentiated tensor (same shape as y), in the epilogue -- fn_prologue_bwd fused */
  int32_t epilogue;
  int32_t adj_prologue;
  int32_t adj_accumulate;
  int32_t reserved;
  const void*  res;
  void*        y_planes;    /* optional second output (FN_EPI_BIAS, tensor cores only): concat_elu(y) in planes form, see fn_conv_desc */
} fn_tconv_desc;

int         fn_abi_version(void);
/* sizeof the descriptor structs as this library was compiled (binding sanity check) */
int         fn_sizeof_conv_desc(void);
int         fn_sizeof_tconv_desc(void);
const char* fn_last_error(void);
/* number of kernel launches issued through this library by the calling process */
uint64_t    fn_launch_count(void);
/* 1 if the last fn_conv_fwd / fn_tconv_fwd on this thread ran a tcgen05 kernel */
int         fn_last_used_tcgen05(void);

/* tf.nn.conv2d + bias_add (+ fused neighbours)          flow_architecture.py:57-68, 129-165 */
int fn_conv_fwd(const fn_conv_desc* d, void* stream);
/* 1 if fn_conv_fwd(d) would run on the tensor cores (pointers are not dereferenced, only tested for NULL), else 0 */
int fn_conv_fwd_has_tcgen05(const fn_conv_desc* d);
/* which tensor-core weight image the kernel that will take this conv reads: 0 = fn_pack_conv_weights_tc of the weights as they are,
 * 1 = pass-major (the two-pass 256-channel kernel, csrc/conv_s1w.cu: the image of K rows {0..127, 256..383} followed by that of
 * {128..255, 384..511}).  w_tc may be NULL in the query.  The ONE place this decision is made: callers pack what it says. */
int fn_conv_weights_layout(const fn_conv_desc* d);
/* tensor-core weight image (as fn_pack_conv_weights_tc, fn_conv_weights_tc_bytes(ksize, K, 0, Cout, 0) bytes) straight from an
 * fp32 variable: logical weights W'[tap][k][n] = src[flip_taps ? T-1-tap : tap][k][n], or src[..][n][k] if transpose.
 * (flip, transpose) = (1, 1) on an HWIO variable gives the B operand of that conv's data gradient. */
int fn_pack_conv_weights_tc_f32(const float* w, void* dst, int32_t ksize, int32_t K, int32_t Cout, int32_t flip_taps,
                                int32_t transpose, void* stream);
/* tf.nn.conv2d_transpose + bias_add                      flow_architecture.py:70-87 */
int fn_tconv_fwd(const fn_tconv_desc* d, void* stream);
/* 1 if fn_tconv_fwd(d) would run on the tensor cores (for FN_EPI_PRO_BWD: if the fused kernel exists), else 0 */
int fn_tconv_fwd_has_tcgen05(const fn_tconv_desc* d);
/* size in bytes of / packer for the tensor-core weight image of a conv (host-side helper
 * that runs a small device kernel): src = bf16 [T, Keff, Cout] (+ optional [Kskip, Cout]) */
int64_t fn_conv_weights_tc_bytes(int32_t ksize, int32_t Keff, int32_t Kskip, int32_t Cout, int32_t gated);
int fn_pack_conv_weights_tc(const void* w, const void* w_skip, void* dst, int32_t ksize,
                            int32_t Keff, int32_t Kskip, int32_t Cout, int32_t gated, void* stream);

/* The whole 8x16 level of the default network in ONE launch (csrc/conv_l5.cu): res_block "resnet_5_downsample"
 * (stride 2), res_block "resnet_5_0" and transpose_conv_layer "up_conv_1" -- flow_architecture.py:207-215 with the
 * reference's flags (nr_res_blocks 1, gated, concat_elu, keep_prob 1) at the 128x256 grid, where an image is one
 * 128-pixel tile at this level and the activations stay in shared memory between the five convolutions.
 *   x: bf16 [B,16,32,64] (output of "resnet_4_0")  ->  y: bf16 [B,16,32,64] (output of "up_conv_1")
 *   w_stream: the tensor-core images (fn_pack_conv_weights_tc) of resnet_5_downsample_conv_1 / _conv_2,
 *             resnet_5_0_conv_1 / _conv_2 and up_conv_1 ([9,128,64] GEMM-B form), back to back: fn_level5_weight_bytes()
 *   bias:     fp32 [128 | 256 | 128 | 256 | 64], the five bias vectors back to back
 * Inference only; any other shape: FN_E_UNSUPPORTED (the layer-by-layer calls above are the general path). */
typedef struct fn_level5_desc {
  int32_t B, H, W, C;       /* input: H = 16, W = 32, C = 64 */
  const void*  x;
  const void*  w_stream;
  const float* bias;
  void*        y;
  void*        y_planes;    /* optional: concat_elu(y) in planes form [B,16,16,32,8] (see fn_conv_desc) */
  /* 1: two CTAs per image (a thread-block cluster, csrc/conv_l5c.cu), each computing half of every convolution's output
   * channels.  w_stream then holds rank 0's five images followed by rank 1's (fn_level5_split_weight_bytes() in all), where
   * rank r's image of a convolution packs its output columns [r F/2, (r+1) F/2) -- for the gated ones the u columns then
   * the matching v columns -- and bias is fp32 [2][64 | 64+64 | 64 | 64+64 | 32] in the same order. */
  int32_t      split;
  int32_t      reserved;
} fn_level5_desc;
int64_t fn_level5_weight_bytes(void);
int64_t fn_level5_split_weight_bytes(void);
int fn_level5_supported(int32_t H, int32_t W, int32_t C);
int fn_level5_fwd(const fn_level5_desc* d, void* stream);

/* ---- the 16x32 level of the 128x256 grid in two launches around fn_level5_fwd (csrc/conv_l4.cu), flow_architecture.py:202-205,
 * 217-222: a cluster of two CTAs per image, activations resident in shared memory between the convolutions.
 *   up = 0: x = x3 bf16 [B,32,64,32] (output of "resnet_3_0") -> y = x4 bf16 [B,16,32,64] (output of "resnet_4_0");
 *           w_stream: tensor-core images of resnet_4_downsample_conv_1 / _conv_2, resnet_4_0_conv_1 / _conv_2, back to back;
 *           bias fp32 [64 | 128 | 64 | 128]
 *   up = 1: x = z bf16 [B,16,32,64] (output of "up_conv_1"), skip = x4 -> y bf16 [B,32,64,32] (output of "up_conv_2");
 *           w_stream: resnet_up_1_0_conv_1 with the nin's 8 k-steps appended (fn_pack_conv_weights_tc with w_skip),
 *           resnet_up_1_0_conv_2, up_conv_2 ([9,64,32] GEMM-B form); bias fp32 [64 (conv_1 + nin) | 128 | 32]
 * Inference only, default flags; any other shape: FN_E_UNSUPPORTED (the layer-by-layer calls are the general path). */
typedef struct fn_level4_desc {
  int32_t B, up;
  const void*  x;
  const void*  skip;        /* up = 1 only */
  const void*  w_stream;    /* fn_level4_weight_bytes(up) bytes */
  const float* bias;
  void*        y;
} fn_level4_desc;
int64_t fn_level4_weight_bytes(int32_t up);
int fn_level4_fwd(const fn_level4_desc* d, void* stream);

/* ---- levels 4 and 5 of the 128x256 grid (twelve convolutions, :202-222) in ONE launch: the stages of fn_level4_fwd(up = 0),
 * fn_level5_fwd(split = 1) and fn_level4_fwd(up = 1) back to back inside one cluster kernel (csrc/conv_l4.cu, MODE_ALL).
 *   x = x3 bf16 [B,32,64,32] -> y bf16 [B,32,64,32] (output of "up_conv_2"); x4 and z: bf16 [B,16,32,64] scratch (the outputs of
 *   "resnet_4_0" and "up_conv_1", valid after the call); weight streams and biases exactly as for the three calls it replaces. */
typedef struct fn_level45_desc {
  int32_t B, reserved;
  const void*  x;
  void*        x4;
  void*        z;
  const void*  w_down;    const float* bias_down;      /* fn_level4_weight_bytes(0), fp32 [384] */
  const void*  w_level5;  const float* bias_level5;    /* fn_level5_split_weight_bytes(), fp32 [2][416] */
  const void*  w_up;      const float* bias_up;        /* fn_level4_weight_bytes(1), fp32 [224] */
  void*        y;
} fn_level45_desc;
int fn_level45_fwd(const fn_level45_desc* d, void* stream);

/* ---- the stride-1 convolutions of the 32x64 level of the 128x256 grid in two launches (csrc/conv_l3.cu), :198-201, 223-228;
 * same organisation as fn_level4_fwd, eight tiles per CTA.
 *   up = 0: x = conv_1 output of "resnet_3_downsample" bf16 [B,32,64,32], aux = x2 bf16 [B,64,128,16] (that block's input,
 *           for the pooled / front-padded shortcut) -> y = x3 bf16 [B,32,64,32] (output of "resnet_3_0");
 *           w_stream: tensor-core images of resnet_3_downsample_conv_2, resnet_3_0_conv_1 / _conv_2; bias fp32 [64 | 32 | 64]
 *   up = 1: x = conv_1 (+ nin) output of "resnet_up_2_0" bf16 [B,32,64,32], aux = z bf16 [B,32,64,32] (the block's input =
 *           output of "up_conv_2") -> y bf16 [B,64,128,16] (output of "up_conv_3");
 *           w_stream: resnet_up_2_0_conv_2, up_conv_3 ([9,32,16] GEMM-B form); bias fp32 [64 | 16]
 * Inference only, default flags. */
typedef struct fn_level3_desc {
  int32_t B, up;
  const void*  x;
  const void*  aux;
  const void*  w_stream;    /* fn_level3_weight_bytes(up) bytes */
  const float* bias;
  void*        y;
} fn_level3_desc;
int64_t fn_level3_weight_bytes(int32_t up);
int fn_level3_fwd(const fn_level3_desc* d, void* stream);

/* standalone nonlinearity (the public concat_elu()/set_nonlinearity() API) :14-29
 * x: bf16 [n, C] -> y: bf16 [n, C or 2C] */
int fn_nonlinearity(const void* x, void* y, int64_t n, int32_t C, int32_t kind, void* stream);
/* dtype plumbing at the boundary: u8/f32 -> bf16 and bf16 -> f32, n elements */
int fn_cast_f32_to_bf16(const float* x, void* y, int64_t n, void* stream);
int fn_cast_u8_to_bf16(const uint8_t* x, void* y, int64_t n, void* stream);
int fn_cast_bf16_to_f32(const void* x, float* y, int64_t n, void* stream);

/* tf.nn.l2_loss(p - t) = sum((p-t)^2)/2                  flow_net.py:39-42
 * p, t: fp32 [n]; loss: fp32 [1] (accumulated into: zero it first); grad (optional): fp32 [n] = p - t */
int fn_l2_loss(const float* p, const float* t, float* loss, float* grad, int64_t n, void* stream);

/* tf.train.AdamOptimizer.apply (TF1 ApplyAdam)           flow_net.py:44-46; SURVEY App. B.10
 * flat fp32 buffers of n elements; step is the 1-based step count */
int fn_adam_step(float* param, const float* grad, float* m, float* v, int64_t n, int64_t step,
                 double lr, double beta1, double beta2, double eps, void* stream);

/* ---- training step: adjoints TF autodiff builds for flow_net.train (model/flow_net.py:44-46).
 * Data gradients of the convolutions are FORWARD launches on the gradient tensor (fn_conv_fwd with
 * flipped/transposed weights for stride 1; fn_tconv_fwd <-> stride-2 fn_conv_fwd for the strided pair);
 * activation gradients are bf16, parameter gradients fp32. */

/* gate y = u*sigmoid(v) (flow_architecture.py:150-151): c2 = [u|v] bf16 [npix,2F], g bf16 [npix,F] -> dc2 bf16 [npix,2F] */
int fn_gate_bwd(const void* c2, const void* g, void* dc2, int64_t npix, int32_t F, void* stream);
/* adjoint of the prologue nonlinearity (+dropout mask of the forward): dA bf16 [npix, mult*C], x bf16 [npix,C]
 * -> dx bf16 [npix,C] (accumulate != 0: added to the existing dx) */
int fn_prologue_bwd(const void* dA, const void* x, void* dx, int64_t npix, int32_t C, int32_t kind, int32_t accumulate,
                    float keep_prob, uint64_t seed, void* stream);
/* the same with a device-resident addend to the seed (see fn_conv_desc.dropout_seed_dev) */
int fn_prologue_bwd_dev(const void* dA, const void* x, void* dx, int64_t npix, int32_t C, int32_t kind, int32_t accumulate,
                        float keep_prob, uint64_t seed, const uint64_t* seed_dev, void* stream);
/* data gradient of the nin skip (flow_architecture.py:136-142) fused with the adjoint of its concat prologue:
 * dA[k] = sum_f dy[f] w[k][f]; da[c] (+)= dA[c] f'(a[c]) - dA[Ca+c] f'(-a[c]).  dy bf16 [npix,F] (F in 8,16,32,64),
 * w fp32 [2 Ca, F] = the '{idx}_fc/weights' variable itself, a bf16 [npix,Ca] -> da bf16 [npix,Ca] */
int fn_nin_bwd(const void* dy, const float* w, const void* a, void* da, int64_t npix, int32_t F, int32_t Ca, int32_t kind,
               int32_t accumulate, void* stream);
/* adjoint of the res_block shortcut (avg-pool 2x2 'SAME' if pooled, front channel pad; :153-165):
 * g bf16 [B,OH,OW,F] -> dx bf16 [B,H,W,Cin] */
int fn_shortcut_bwd(const void* g, void* dx, int32_t B, int32_t OH, int32_t OW, int32_t F, int32_t H, int32_t W,
                    int32_t Cin, int32_t pooled, int32_t accumulate, void* stream);
/* loss += sum((p-t)^2)/2 (flow_net.py:39-42) and the gradient w.r.t. the pre-tanh tail output, written as
 * bf16 [npix, 8] with channels 2..7 zero (so the tail's dgrad/wgrad run as 8-channel launches) */
int fn_tanh_l2_bwd(const float* p, const float* t, float* loss, void* dpre, int64_t npix, void* stream);
/* db[n] = sum_pix dY[pix][n]; workspace: fn_bias_grad_workspace() floats */
int64_t fn_bias_grad_workspace(int64_t npix, int32_t N);
int fn_bias_grad(const void* dY, float* db, float* workspace, int64_t npix, int32_t N, void* stream);
/* dW[tap][k][n] = sum_pix pro(x)[pix*s+tap-pad][k] * dY[pix][n]; d describes the FORWARD launch (x, dims, ksize,
 * stride, prologue, keep_prob, dropout_seed, impl); dW fp32 [k*k, Keff, Cout]; workspace: fn_conv_wgrad_workspace() floats.
 * FN_IMPL_AUTO runs the tcgen05 kernel (contraction over pixels, MN-major operands) when Cin, Cout are multiples of 8. */
int64_t fn_conv_wgrad_workspace(const fn_conv_desc* d);
int fn_conv_wgrad(const fn_conv_desc* d, const void* dY, float* dW, float* workspace, void* stream);
/* the same plus db[n] = sum_pix dY[pix][n] (fp32 [Cout]); on the tcgen05 path the bias gradient is one more accumulator of
 * the same kernel, fed by a plane of ones */
int fn_conv_wgrad_bias(const fn_conv_desc* d, const void* dY, float* dW, float* db, float* workspace, void* stream);
/* transposed conv: dW fp32 [9, Cout, Cin] (the variable's own [k,k,Cout,Cin] layout, flow_architecture.py:74) from
 * x bf16 [B,H,W,Cin] and dY bf16 [B,2H,2W,Cout].  d->impl selects the kernel as in the forward calls. */
int64_t fn_tconv_wgrad_workspace(const fn_tconv_desc* d);
int fn_tconv_wgrad(const fn_tconv_desc* d, const void* dY, float* dW, float* workspace, void* stream);

/* ---- dataset files (host code, no GPU): the TFRecord / tf.train.Example format written by
 * utils/createTFRecords.py:45-56 and read by input/flow_input.py:14-30 of the reference ------------------------- */
/* walks a whole TFRecord file image; returns the record count (or FN_E_INVAL on a truncated / corrupt file) and
 * fills payload offsets / lengths for the first max_records records (both arrays may be NULL to just count).
 * check_crc != 0 verifies the masked CRC32-C of every length word and payload. */
int64_t fn_tfrecord_index(const uint8_t* buf, int64_t nbytes, int64_t* offsets, int64_t* lengths, int64_t max_records,
                          int32_t check_crc);
/* one Example payload -> 'boundary' (uint8, boundary_bytes = H*W) and 'sflow' (float32, sflow_bytes = H*W*2*4);
 * FN_E_INVAL when a feature is missing or has another size */
int fn_tfrecord_parse_flow_example(const uint8_t* payload, int64_t len, uint8_t* boundary, int64_t boundary_bytes, float* sflow,
                                   int64_t sflow_bytes);
/* n records (payload offsets / lengths into one file image) -> consecutive slots of a batch, on up to nthreads host threads */
int fn_tfrecord_parse_flow_batch(const uint8_t* buf, const int64_t* offsets, const int64_t* lengths, int64_t n, uint8_t* boundary,
                                 int64_t boundary_bytes, float* sflow, int64_t sflow_bytes, int32_t nthreads);
/* writes one framed record holding such an Example; dst == NULL returns the size needed; returns bytes written */
int64_t fn_tfrecord_write_flow_example(uint8_t* dst, int64_t capacity, const uint8_t* boundary, int64_t boundary_bytes,
                                       const float* sflow, int64_t sflow_bytes);
uint32_t fn_crc32c(const uint8_t* p, int64_t n);          /* SSE4.2 instruction when the CPU has it */
uint32_t fn_crc32c_portable(const uint8_t* p, int64_t n); /* table-driven */

/* ---- labelled-data generator: D2Q9 BGK lattice-Boltzmann channel flow (replaces the MechSys program
 * fluid_generator/generate_xiao_fluid_flow.cpp: Zou-He parabolic inlet :147-165, density outlet :170-187, bounce-back
 * solids, tau = 3 nu + 1/2 with nu = u_max*200/Re :222-231, start at rho0 = 1, v0 = (0.08, 0) :355-363, 10000 steps) -- */
typedef struct fn_lbm_desc {
  int32_t B, nx, ny;        /* B independent simulations on nx x ny lattices (reference: nx = 2*ny + 128 = 384, ny = 128) */
  int32_t dtype;            /* 0: fp32 populations, 1: fp64 (MechSys computes in double) */
  double  tau;              /* BGK relaxation time (reference: 0.65) */
  double  u_max;            /* peak of the parabolic inlet profile (reference: 0.1) */
  double  rho_out;          /* outlet density (reference: 1.0) */
  const uint8_t* solid;     /* [B, ny, nx], 1 = solid */
  void*   f_a;              /* populations [B, 9, ny, nx] (current state) */
  void*   f_b;              /* second lattice of the same size */
} fn_lbm_desc;
int fn_lbm_init(const fn_lbm_desc* d, double rho0, double vx0, double vy0, void* stream);   /* f_a <- equilibrium */
int fn_lbm_steps(const fn_lbm_desc* d, int32_t nsteps, void* stream);                       /* nsteps even; result in f_a */
/* vel: fp32 [B, ny, nx, 2] (and rho: fp32 [B, ny, nx] or NULL) of the current state; solid cells report 0 */
int fn_lbm_fields(const fn_lbm_desc* d, float* vel, float* rho, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FLOWNET_H_ */
