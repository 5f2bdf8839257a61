// extern "C" surface of libflownet.so (declared in include/flownet.h): argument validation,
// error plumbing and dispatch to the tcgen05 (conv_tc.cu) or CUDA-core (conv_simt.cu) kernels.
#include <atomic>
#include <cstdarg>
#include <cstring>

#include <stdlib.h>

#include "common.cuh"
#include "tma.cuh"

namespace fn {

static thread_local char g_err[512] = "";
static thread_local int g_last_tc = 0;
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
bool pdl_enabled() {
  static const bool v = [] {
    const char* e = getenv("FLOWNET_PDL");      // measured on B200: 97.2k vs 97.5k img/s -- the graph already runs the
    return e && e[0] == '1';                     // kernels back to back, so it stays opt-in
  }();
  return v;
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
void note_tcgen05(bool used) { g_last_tc = used ? 1 : 0; }
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

// implemented in the other translation units
int conv_fwd_simt(const fn_conv_desc& d, cudaStream_t stream);
int tconv_fwd_simt(const fn_tconv_desc& d, cudaStream_t stream);
bool conv_tc_supported(const fn_conv_desc& d, const char** why);
int conv_fwd_tc(const fn_conv_desc& d, cudaStream_t stream);
bool tconv_tc_supported(const fn_tconv_desc& d, const char** why);
int tconv_fwd_tc(const fn_tconv_desc& d, cudaStream_t stream);
long long conv_weights_tc_bytes(int ksize, int Keff, int Kskip, int Cout, int gated);
int pack_conv_weights_tc_f32(const float* w, void* dst, int ksize, int K, int Cout, int flip, int transpose, cudaStream_t stream);
int pack_conv_weights_tc(const void* w, const void* w_skip, void* dst, int ksize, int Keff, int Kskip, int Cout,
                         int gated, cudaStream_t stream);
int nonlinearity(const void* x, void* y, long long n, int C, int kind, cudaStream_t s);
template <typename TI, typename TO>
int cast(const TI* x, TO* y, long long n, cudaStream_t s);
int l2_loss(const float* p, const float* t, float* loss, float* grad, long long n, cudaStream_t s);
int adam_step(float* param, const float* grad, float* m, float* v, long long n, long long step, double lr, double b1,
              double b2, double eps, cudaStream_t s);
int level5_fwd(const fn_level5_desc& d, cudaStream_t stream);      // conv_l5.cu
long long level5_weight_bytes();
int level5c_fwd(const fn_level5_desc& d, cudaStream_t stream);     // conv_l5c.cu
long long level5c_weight_bytes();
int level4_fwd(const fn_level4_desc& d, cudaStream_t stream);       // conv_l4.cu
long long level4_weight_bytes(int up);
int level45_fwd(const fn_level45_desc& d, cudaStream_t stream);     // conv_l4.cu
bool conv_tail_supported(const fn_conv_desc& d);                    // conv_tail.cu
bool conv_weights_pass_major(const fn_conv_desc& d);                // conv_tc.cu
int conv_tail_fwd(const fn_conv_desc& d, cudaStream_t stream);
int level3_fwd(const fn_level3_desc& d, cudaStream_t stream);       // conv_l3.cu
long long level3_weight_bytes(int up);

static int validate_conv(const fn_conv_desc& d) {
  FN_REQUIRE(d.B > 0 && d.H > 0 && d.W > 0 && d.Cin > 0 && d.Cout > 0, FN_E_INVAL,
             "fn_conv_fwd: non-positive dimension (B=%d H=%d W=%d Cin=%d Cout=%d)", d.B, d.H, d.W, d.Cin, d.Cout);
  FN_REQUIRE(d.x && d.w && d.bias && (d.y || d.y_planes), FN_E_INVAL, "fn_conv_fwd: null x/w/bias/y");
  FN_REQUIRE(aligned16(d.y_planes), FN_E_ALIGN, "fn_conv_fwd: y_planes must be 16-byte aligned");
  if (d.x_planes || d.skip_planes)
    FN_REQUIRE(d.x_planes && (!d.skip || d.skip_planes) && d.prologue == FN_PRO_CONCAT_ELU && d.impl != FN_IMPL_SIMT, FN_E_UNSUPPORTED,
               "fn_conv_fwd: planes operands need x_planes (and skip_planes with a skip), the concat_elu prologue, tcgen05");
  if (d.x_u8)
    FN_REQUIRE(d.Cin == 1 && d.Cout == 8 && d.ksize == 3 && d.stride == 1 && d.prologue == FN_PRO_NONE && d.epilogue == FN_EPI_BIAS && !d.skip &&
                   !d.x_planes && d.keep_prob >= 1.0f,
               FN_E_UNSUPPORTED, "fn_conv_fwd: a uint8 input is read by the 1 -> 8 channel head conv only");
  if (d.res_u8)
    FN_REQUIRE(d.res && d.Cres == 1 && !d.res_pool && d.epilogue == FN_EPI_GATE_RES && d.impl != FN_IMPL_SIMT, FN_E_UNSUPPORTED,
               "fn_conv_fwd: a uint8 shortcut is the one-channel boundary image under a gated block (tcgen05)");
  FN_REQUIRE(d.ksize == 3 || d.ksize == 1, FN_E_UNSUPPORTED, "fn_conv_fwd: kernel size %d (only 1 and 3)", d.ksize);
  FN_REQUIRE(d.stride == 1 || d.stride == 2, FN_E_UNSUPPORTED, "fn_conv_fwd: stride %d (only 1 and 2)", d.stride);
  FN_REQUIRE(d.prologue >= FN_PRO_NONE && d.prologue <= FN_PRO_RELU, FN_E_INVAL, "fn_conv_fwd: bad prologue %d",
             d.prologue);
  FN_REQUIRE(d.epilogue >= FN_EPI_BIAS && d.epilogue <= FN_EPI_PRO_BWD, FN_E_INVAL, "fn_conv_fwd: bad epilogue %d",
             d.epilogue);
  FN_REQUIRE(d.impl >= FN_IMPL_AUTO && d.impl <= FN_IMPL_TCGEN05, FN_E_INVAL, "fn_conv_fwd: bad impl %d", d.impl);
  FN_REQUIRE(d.Cin == 1 || (d.Cin % 8) == 0, FN_E_ALIGN, "fn_conv_fwd: Cin=%d must be 1 or a multiple of 8", d.Cin);
  FN_REQUIRE(aligned16(d.x) && aligned16(d.y) && aligned16(d.w), FN_E_ALIGN, "fn_conv_fwd: x/y/w must be 16-byte aligned");
  if (d.epilogue == FN_EPI_GATE_BWD)
    FN_REQUIRE(d.res && d.Cres == d.Cout / 2 && d.RH == d.H && d.RW == d.W && !d.res_pool && d.stride == 1 && !d.skip &&
                   d.impl != FN_IMPL_SIMT,
               FN_E_UNSUPPORTED, "fn_conv_fwd: gate-backward epilogue needs res = gradient [B,H,W,Cout/2], stride 1, no skip, tcgen05");
  if (d.epilogue == FN_EPI_PRO_BWD)
    FN_REQUIRE(d.res && d.adj_prologue >= FN_PRO_NONE && d.adj_prologue <= FN_PRO_RELU && d.impl != FN_IMPL_SIMT, FN_E_UNSUPPORTED,
               "fn_conv_fwd: prologue-adjoint epilogue needs res = the differentiated input, a valid adj_prologue, tcgen05");
  if (d.epilogue == FN_EPI_GATE_RES || d.epilogue == FN_EPI_GATE_BWD)
    FN_REQUIRE((d.Cout % 16) == 0, FN_E_ALIGN, "fn_conv_fwd: gated Cout=%d must be a multiple of 16", d.Cout);
  else if (d.epilogue != FN_EPI_TANH)
    FN_REQUIRE((d.Cout % 8) == 0, FN_E_ALIGN, "fn_conv_fwd: Cout=%d must be a multiple of 8", d.Cout);
  if (d.skip) {
    FN_REQUIRE(d.w_skip && d.Cskip > 0 && (d.Cskip % 8) == 0 && d.SH > 0 && d.SW > 0 && aligned16(d.skip), FN_E_INVAL,
               "fn_conv_fwd: bad skip operand (Cskip=%d SH=%d SW=%d)", d.Cskip, d.SH, d.SW);
    const int OH = (d.H + d.stride - 1) / d.stride, OW = (d.W + d.stride - 1) / d.stride;
    FN_REQUIRE(d.SH <= OH && d.SW <= OW, FN_E_INVAL, "fn_conv_fwd: skip %dx%d larger than output %dx%d", d.SH, d.SW,
               OH, OW);
  }
  if (d.epilogue == FN_EPI_GATE_RES || d.epilogue == FN_EPI_RES) {
    const int F = d.epilogue == FN_EPI_GATE_RES ? d.Cout / 2 : d.Cout;
    if (d.res) {
      FN_REQUIRE(d.Cres > 0 && d.Cres <= F && d.RH > 0 && d.RW > 0, FN_E_INVAL,
                 "fn_conv_fwd: bad residual operand (Cres=%d F=%d RH=%d RW=%d)", d.Cres, F, d.RH, d.RW);
    }
  }
  return 0;
}

}  // namespace fn

using namespace fn;

extern "C" {

int fn_abi_version(void) { return FN_ABI_VERSION; }
int fn_sizeof_conv_desc(void) { return (int)sizeof(fn_conv_desc); }
int fn_sizeof_tconv_desc(void) { return (int)sizeof(fn_tconv_desc); }
const char* fn_last_error(void) { return g_err; }
uint64_t fn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
int fn_last_used_tcgen05(void) { return g_last_tc; }

int fn_conv_fwd(const fn_conv_desc* d, void* stream) {
  FN_REQUIRE(d != nullptr, FN_E_INVAL, "fn_conv_fwd: null descriptor");
  int rc = validate_conv(*d);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const char* why = "";
  if (d->impl != FN_IMPL_SIMT && conv_tail_supported(*d)) return conv_tail_fwd(*d, s);
  if (d->impl != FN_IMPL_SIMT && conv_tc_supported(*d, &why)) {
    note_tcgen05(true);
    return conv_fwd_tc(*d, s);
  }
  FN_REQUIRE(d->impl != FN_IMPL_TCGEN05 && d->epilogue != FN_EPI_GATE_BWD && d->epilogue != FN_EPI_PRO_BWD && !d->x_planes &&
                 !d->skip_planes,
             FN_E_UNSUPPORTED, "fn_conv_fwd: no tcgen05 kernel for this shape: %s", why);
  if (d->y_planes || !d->y)     // CUDA cores: only the 1-channel head writes planes
    FN_REQUIRE(d->Cin == 1 && (d->Cout == 8 || d->Cout == 16) && d->ksize == 3 && d->stride == 1 && d->epilogue == FN_EPI_BIAS &&
                   d->prologue == FN_PRO_NONE && !d->skip && d->keep_prob >= 1.0f,
               FN_E_UNSUPPORTED, "fn_conv_fwd: no kernel writes a planes output for this shape: %s", why);
  note_tcgen05(false);
  return conv_fwd_simt(*d, s);
}

int fn_pack_conv_weights_tc_f32(const float* w, void* dst, int32_t ksize, int32_t K, int32_t Cout, int32_t flip_taps,
                                int32_t transpose, void* stream) {
  FN_REQUIRE(w && dst && (ksize == 1 || ksize == 3) && K > 0 && Cout > 0, FN_E_INVAL, "fn_pack_conv_weights_tc_f32: bad arguments");
  return pack_conv_weights_tc_f32(w, dst, ksize, K, Cout, flip_taps, transpose, (cudaStream_t)stream);
}

int fn_conv_fwd_has_tcgen05(const fn_conv_desc* d) {
  if (d == nullptr || validate_conv(*d) != 0 || d->impl == FN_IMPL_SIMT) return 0;
  const char* why = "";
  return (conv_tail_supported(*d) || conv_tc_supported(*d, &why)) ? 1 : 0;
}

int fn_conv_weights_layout(const fn_conv_desc* d) {
  if (d == nullptr || d->impl == FN_IMPL_SIMT || d->skip != nullptr || d->x_planes) return 0;
  fn_conv_desc q = *d;
  if (q.w_tc == nullptr) q.w_tc = q.w;          // any non-null pointer: the query launches nothing
  if (validate_conv(q) != 0) return 0;
  return conv_weights_pass_major(q) ? 1 : 0;
}

static int validate_tconv(const fn_tconv_desc* d) {
  FN_REQUIRE(d != nullptr, FN_E_INVAL, "fn_tconv_fwd: null descriptor");
  FN_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, FN_E_INVAL,
             "fn_tconv_fwd: non-positive dimension");
  FN_REQUIRE(d->x && d->w && d->bias && d->y, FN_E_INVAL, "fn_tconv_fwd: null x/w/bias/y");
  FN_REQUIRE(aligned16(d->y_planes) && !(d->y_planes && (d->epilogue != FN_EPI_BIAS || d->impl == FN_IMPL_SIMT)), FN_E_UNSUPPORTED,
             "fn_tconv_fwd: y_planes needs 16-byte alignment, the bias epilogue and a tensor-core kernel");
  FN_REQUIRE((d->Cin % 8) == 0 && (d->Cout % 8) == 0, FN_E_ALIGN, "fn_tconv_fwd: Cin=%d Cout=%d must be multiples of 8",
             d->Cin, d->Cout);
  FN_REQUIRE(aligned16(d->x) && aligned16(d->y) && aligned16(d->w), FN_E_ALIGN, "fn_tconv_fwd: x/y/w must be 16-byte aligned");
  FN_REQUIRE(d->impl >= FN_IMPL_AUTO && d->impl <= FN_IMPL_TCGEN05, FN_E_INVAL, "fn_tconv_fwd: bad impl %d", d->impl);
  FN_REQUIRE(d->epilogue == FN_EPI_BIAS || d->epilogue == FN_EPI_PRO_BWD, FN_E_INVAL, "fn_tconv_fwd: bad epilogue %d", d->epilogue);
  if (d->epilogue == FN_EPI_PRO_BWD)
    FN_REQUIRE(d->res && aligned16(d->res) && d->Cout == d->Cin && d->impl != FN_IMPL_SIMT &&
                   (d->adj_prologue == FN_PRO_CONCAT_ELU || d->adj_prologue == FN_PRO_CONCAT_RELU),
               FN_E_UNSUPPORTED, "fn_tconv_fwd: prologue-adjoint epilogue needs res, Cout == Cin, a concat prologue, tcgen05");
  return 0;
}

int fn_tconv_fwd(const fn_tconv_desc* d, void* stream) {
  int rc = validate_tconv(d);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const char* why = "";
  if (d->impl != FN_IMPL_SIMT && tconv_tc_supported(*d, &why)) {
    note_tcgen05(true);
    return tconv_fwd_tc(*d, s);
  }
  FN_REQUIRE(d->impl != FN_IMPL_TCGEN05 && d->epilogue != FN_EPI_PRO_BWD && !d->y_planes, FN_E_UNSUPPORTED,
             "fn_tconv_fwd: no tcgen05 kernel for this shape: %s", why);
  note_tcgen05(false);
  return tconv_fwd_simt(*d, s);
}

int fn_tconv_fwd_has_tcgen05(const fn_tconv_desc* d) {
  if (d == nullptr || d->impl == FN_IMPL_SIMT) return 0;
  const int quiet = validate_tconv(d);
  if (quiet != 0) return 0;
  const char* why = "";
  return tconv_tc_supported(*d, &why) ? 1 : 0;
}

int64_t fn_conv_weights_tc_bytes(int32_t ksize, int32_t Keff, int32_t Kskip, int32_t Cout, int32_t gated) {
  return conv_weights_tc_bytes(ksize, Keff, Kskip, Cout, gated);
}

int fn_pack_conv_weights_tc(const void* w, const void* w_skip, void* dst, int32_t ksize, int32_t Keff, int32_t Kskip,
                            int32_t Cout, int32_t gated, void* stream) {
  FN_REQUIRE(w && dst, FN_E_INVAL, "fn_pack_conv_weights_tc: null w/dst");
  return pack_conv_weights_tc(w, w_skip, dst, ksize, Keff, Kskip, Cout, gated, (cudaStream_t)stream);
}

int64_t fn_level5_weight_bytes(void) { return level5_weight_bytes(); }
int64_t fn_level5_split_weight_bytes(void) { return level5c_weight_bytes(); }
int fn_level5_supported(int32_t H, int32_t W, int32_t C) { return (H == 16 && W == 32 && C == 64 && tma::get_encode() != nullptr) ? 1 : 0; }
int fn_level5_fwd(const fn_level5_desc* d, void* stream) {
  FN_REQUIRE(d != nullptr, FN_E_INVAL, "fn_level5_fwd: null descriptor");
  FN_REQUIRE(d->B > 0 && d->x && d->w_stream && d->bias && d->y, FN_E_INVAL, "fn_level5_fwd: null pointer or B = %d", d->B);
  FN_REQUIRE(d->H == 16 && d->W == 32 && d->C == 64, FN_E_UNSUPPORTED,
             "fn_level5_fwd: built for the 128x256 grid's level-5 input [B,16,32,64], got [B,%d,%d,%d]", d->H, d->W, d->C);
  FN_REQUIRE(aligned16(d->x) && aligned16(d->y) && aligned16(d->w_stream), FN_E_ALIGN, "fn_level5_fwd: x/y/w_stream must be 16-byte aligned");
  return d->split ? level5c_fwd(*d, (cudaStream_t)stream) : level5_fwd(*d, (cudaStream_t)stream);
}

int64_t fn_level4_weight_bytes(int32_t up) { return level4_weight_bytes(up); }
int fn_level4_fwd(const fn_level4_desc* d, void* stream) {
  FN_REQUIRE(d != nullptr, FN_E_INVAL, "fn_level4_fwd: null descriptor");
  FN_REQUIRE(d->B > 0 && d->x && d->w_stream && d->bias && d->y && (!d->up || d->skip), FN_E_INVAL, "fn_level4_fwd: null pointer or B = %d", d->B);
  FN_REQUIRE(aligned16(d->x) && aligned16(d->y) && aligned16(d->w_stream) && aligned16(d->skip), FN_E_ALIGN,
             "fn_level4_fwd: x/skip/y/w_stream must be 16-byte aligned");
  return level4_fwd(*d, (cudaStream_t)stream);
}

int fn_level45_fwd(const fn_level45_desc* d, void* stream) {
  FN_REQUIRE(d != nullptr, FN_E_INVAL, "fn_level45_fwd: null descriptor");
  FN_REQUIRE(d->B > 0 && d->x && d->x4 && d->z && d->y && d->w_down && d->bias_down && d->w_level5 && d->bias_level5 && d->w_up && d->bias_up,
             FN_E_INVAL, "fn_level45_fwd: null pointer or B = %d", d->B);
  FN_REQUIRE(aligned16(d->x) && aligned16(d->x4) && aligned16(d->z) && aligned16(d->y) && aligned16(d->w_down) && aligned16(d->w_level5) &&
                 aligned16(d->w_up), FN_E_ALIGN, "fn_level45_fwd: tensors and weight streams must be 16-byte aligned");
  FN_REQUIRE(tma::get_encode() != nullptr, FN_E_UNSUPPORTED, "fn_level45_fwd: no tensor-map encoder (driver too old)");
  return level45_fwd(*d, (cudaStream_t)stream);
}
int64_t fn_level3_weight_bytes(int32_t up) { return level3_weight_bytes(up); }
int fn_level3_fwd(const fn_level3_desc* d, void* stream) {
  FN_REQUIRE(d != nullptr, FN_E_INVAL, "fn_level3_fwd: null descriptor");
  FN_REQUIRE(d->B > 0 && d->x && d->aux && d->w_stream && d->bias && d->y, FN_E_INVAL, "fn_level3_fwd: null pointer or B = %d", d->B);
  FN_REQUIRE(aligned16(d->x) && aligned16(d->aux) && aligned16(d->y) && aligned16(d->w_stream), FN_E_ALIGN,
             "fn_level3_fwd: x/aux/y/w_stream must be 16-byte aligned");
  return level3_fwd(*d, (cudaStream_t)stream);
}

int fn_nonlinearity(const void* x, void* y, int64_t n, int32_t C, int32_t kind, void* stream) {
  FN_REQUIRE(x && y && n >= 0 && C > 0, FN_E_INVAL, "fn_nonlinearity: bad arguments");
  FN_REQUIRE(kind >= FN_PRO_CONCAT_ELU && kind <= FN_PRO_RELU, FN_E_INVAL, "fn_nonlinearity: bad kind %d", kind);
  return nonlinearity(x, y, n, C, kind, (cudaStream_t)stream);
}

int fn_cast_f32_to_bf16(const float* x, void* y, int64_t n, void* stream) {
  FN_REQUIRE(x && y && n >= 0, FN_E_INVAL, "fn_cast_f32_to_bf16: bad arguments");
  return cast<float, __nv_bfloat16>(x, (__nv_bfloat16*)y, n, (cudaStream_t)stream);
}
int fn_cast_u8_to_bf16(const uint8_t* x, void* y, int64_t n, void* stream) {
  FN_REQUIRE(x && y && n >= 0, FN_E_INVAL, "fn_cast_u8_to_bf16: bad arguments");
  return cast<uint8_t, __nv_bfloat16>(x, (__nv_bfloat16*)y, n, (cudaStream_t)stream);
}
int fn_cast_bf16_to_f32(const void* x, float* y, int64_t n, void* stream) {
  FN_REQUIRE(x && y && n >= 0, FN_E_INVAL, "fn_cast_bf16_to_f32: bad arguments");
  return cast<__nv_bfloat16, float>((const __nv_bfloat16*)x, y, n, (cudaStream_t)stream);
}

int fn_l2_loss(const float* p, const float* t, float* loss, float* grad, int64_t n, void* stream) {
  FN_REQUIRE(loss && n >= 0 && ((p && t) || n == 0), FN_E_INVAL, "fn_l2_loss: bad arguments");
  return l2_loss(p, t, loss, grad, n, (cudaStream_t)stream);
}

int fn_adam_step(float* param, const float* grad, float* m, float* v, int64_t n, int64_t step, double lr, double beta1,
                 double beta2, double eps, void* stream) {
  FN_REQUIRE(param && grad && m && v && n >= 0 && step >= 1, FN_E_INVAL, "fn_adam_step: bad arguments");
  return adam_step(param, grad, m, v, n, step, lr, beta1, beta2, eps, (cudaStream_t)stream);
}

}  // extern "C"
