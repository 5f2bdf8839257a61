// Shared device/host helpers for libflownet.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <utility>

#include "flownet.h"

namespace fn {

// ---- error plumbing (thread-local message, no exceptions across the ABI)
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
void note_tcgen05(bool used);
int check_launch(const char* what);

#define FN_REQUIRE(cond, code, ...)        \
  do {                                     \
    if (!(cond)) {                         \
      fn::set_error(__VA_ARGS__);          \
      return (code);                       \
    }                                      \
  } while (0)

// ---- programmatic dependent launch: a kernel launched with launch_pdl() may become resident while its predecessor in
// the stream is still running; everything it does before pdl_wait() (barrier init, TMEM allocation) overlaps the
// predecessor's tail, and nothing before pdl_wait() may touch global memory.  pdl_launch_dependents() comes AFTER the
// wait, so at most one kernel runs ahead.  Both instructions are no-ops in a kernel launched the ordinary way.
bool pdl_enabled();        // env FLOWNET_PDL=1 (default off: no measurable gain inside the CUDA graph)
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}
#endif

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// TF 'SAME' padding-before for kernel k, stride s on an axis of n
__host__ __device__ inline int same_pad_before(int n, int k, int s) {
  int out = (n + s - 1) / s;
  int tot = (out - 1) * s + k - n;
  if (tot < 0) tot = 0;
  return tot / 2;
}

// ---- nonlinearities (flow_architecture.py:14-29)
__device__ __forceinline__ float elu_f(float z) { return z > 0.f ? z : (__expf(z) - 1.f); }
__device__ __forceinline__ float sigmoid_f(float z) { return 1.f / (1.f + __expf(-z)); }

// [elu(x) | elu(-x)] of two packed bf16 values (flow_architecture.py:14-17).  t = exp(-|v|) - 1 is computed in fp32 and
// rounded once, exactly as the scalar form (v > 0 ? v : t, v > 0 ? t : -v); the selects are packed maxima: t >= min(v, -v)
// and t <= 0 <= max(v, -v) hold after rounding too (rounding is monotonic and v is a bf16 value), so the results are
// the same values (up to the sign of a zero) for 5.5 instead of 9.5 instructions per element.
__device__ __forceinline__ void concat_elu2(uint32_t xw, uint32_t& o0, uint32_t& o1) {
  const float lo = __uint_as_float(xw << 16), hi = __uint_as_float(xw & 0xffff0000u);
  const float tl = __expf(-fabsf(lo)) - 1.f, th = __expf(-fabsf(hi)) - 1.f;
  const __nv_bfloat162 t2 = __floats2bfloat162_rn(tl, th);
  const __nv_bfloat162 x2 = *reinterpret_cast<const __nv_bfloat162*>(&xw);
  const __nv_bfloat162 a = __hmax2(x2, t2), b = __hmax2(__hneg2(x2), t2);
  o0 = *reinterpret_cast<const uint32_t*>(&a);
  o1 = *reinterpret_cast<const uint32_t*>(&b);
}
__device__ __forceinline__ void concat_elu8(const uint4& in, uint4& o0, uint4& o1) {
  concat_elu2(in.x, o0.x, o1.x);
  concat_elu2(in.y, o0.y, o1.y);
  concat_elu2(in.z, o0.z, o1.z);
  concat_elu2(in.w, o0.w, o1.w);
}

// number of GEMM-K channels produced per raw channel by a prologue
__host__ __device__ inline int pro_mult(int pro) {
  return (pro == FN_PRO_CONCAT_ELU || pro == FN_PRO_CONCAT_RELU) ? 2 : 1;
}

// first (and for concat variants second) prologue output of one raw value
__device__ __forceinline__ void pro_apply(int pro, float v, float& p0, float& p1) {
  switch (pro) {
    case FN_PRO_CONCAT_ELU: {   // one exponential serves both halves: t = exp(-|v|) - 1
      const float t = __expf(-fabsf(v)) - 1.f;
      p0 = v > 0.f ? v : t;
      p1 = v > 0.f ? t : -v;
      break;
    }
    case FN_PRO_ELU:         p0 = elu_f(v);       p1 = 0.f;             break;
    case FN_PRO_CONCAT_RELU: p0 = fmaxf(v, 0.f);  p1 = fmaxf(-v, 0.f);  break;
    case FN_PRO_RELU:        p0 = fmaxf(v, 0.f);  p1 = 0.f;             break;
    default:                 p0 = v;              p1 = 0.f;             break;
  }
}

// ---- counter-based dropout mask (tf.nn.dropout, flow_architecture.py:144-145):
// keep element `idx` iff u(seed, idx) < keep_prob; kept values are scaled by 1/keep_prob.
__host__ __device__ inline uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// One hash serves 4 elements (16 bits each), so an aligned group of 8 consecutive elements -- what every vector
// kernel handles at a time -- costs three mix64 instead of sixteen.  Element idx: group q = idx/8, j = idx%8;
// h1 = mix64(seed ^ mix64(q)), h2 = mix64(h1); u = 16-bit field j%4 of (j < 4 ? h1 : h2); keep iff u < keep_prob*65536.
__host__ __device__ inline uint32_t dropout_threshold(float keep_prob) { return (uint32_t)(keep_prob * 65536.0f); }
__host__ __device__ inline float dropout_scale(uint64_t seed, uint64_t idx, float keep_prob) {
  uint64_t h = mix64(seed ^ mix64(idx >> 3));
  if (idx & 4) h = mix64(h);
  const uint32_t u = (uint32_t)(h >> (16 * (idx & 3))) & 0xFFFFu;
  return u < dropout_threshold(keep_prob) ? 1.0f / keep_prob : 0.0f;
}
// the 8 scales of the aligned group starting at element idx8 (idx8 % 8 == 0)
__host__ __device__ inline void dropout_scale8(uint64_t seed, uint64_t idx8, float keep_prob, float (&sc)[8]) {
  const uint64_t h1 = mix64(seed ^ mix64(idx8 >> 3));
  const uint64_t h2 = mix64(h1);
  const uint32_t thr = dropout_threshold(keep_prob);
  const float inv = 1.0f / keep_prob;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sc[j] = ((uint32_t)(h1 >> (16 * j)) & 0xFFFFu) < thr ? inv : 0.0f;
    sc[4 + j] = ((uint32_t)(h2 >> (16 * j)) & 0xFFFFu) < thr ? inv : 0.0f;
  }
}

union bf16x8 {
  uint4 v;
  __nv_bfloat16 h[8];
  __nv_bfloat162 h2[4];
};

#ifdef __CUDACC__
// ---- generic shortcut of a res_block (flow_architecture.py:153-165): v[0..8) += sc(res)[b, oy, ox, c0 .. c0+8) where sc is
// the optional 2x2/2 'SAME' average pool followed by the FRONT channel pad from Cres to F channels.  The common case
// (pooled, whole 8-channel groups, all four source pixels inside) is inline; everything else (ragged edges, odd channel
// counts, un-pooled channel pads) is one out-of-line scalar routine, so that the tensor-core kernels do not carry a copy
// of it per unrolled epilogue item -- their start-up is bound by instruction fetch.
static __device__ __noinline__ void shortcut_add8_slow(const __nv_bfloat16* res, int Cres, int RH, int RW, int res_pool, int F, int b,
                                                int oy, int ox, int c0, float* v) {
  const int shift = F - Cres;
#pragma unroll 1
  for (int j = 0; j < 8; ++j) {
    const int cr = c0 + j - shift;
    if (cr < 0) continue;
    if (res_pool) {
      float s = 0.f;
      int cnt = 0;
      for (int a = 0; a < 2; ++a)
        for (int c = 0; c < 2; ++c) {
          const int ry = 2 * oy + a, rx = 2 * ox + c;
          if (ry < RH && rx < RW) {
            s += __bfloat162float(res[(((size_t)b * RH + ry) * RW + rx) * Cres + cr]);
            ++cnt;
          }
        }
      v[j] += s / (float)cnt;
    } else {
      v[j] += __bfloat162float(res[(((size_t)b * RH + oy) * RW + ox) * Cres + cr]);
    }
  }
}
__device__ __forceinline__ void shortcut_add8(const __nv_bfloat16* res, int Cres, int RH, int RW, int res_pool, int F, int b, int oy,
                                              int ox, int c0, float (&v)[8]) {
  const int shift = F - Cres;
  if (res_pool && (Cres & 7) == 0 && ((c0 - shift) & 7) == 0 && 2 * oy + 1 < RH && 2 * ox + 1 < RW) {
    if (c0 - shift < 0) return;             // front padding: nothing to add
    const __nv_bfloat16* r00 = res + (((size_t)b * RH + 2 * oy) * RW + 2 * ox) * Cres + (c0 - shift);
    bf16x8 q[4];
    q[0].v = __ldg(reinterpret_cast<const uint4*>(r00));
    q[1].v = __ldg(reinterpret_cast<const uint4*>(r00 + Cres));
    q[2].v = __ldg(reinterpret_cast<const uint4*>(r00 + (size_t)RW * Cres));
    q[3].v = __ldg(reinterpret_cast<const uint4*>(r00 + (size_t)RW * Cres + Cres));
#pragma unroll
    for (int j = 0; j < 8; ++j)
      v[j] += 0.25f * (__bfloat162float(q[0].h[j]) + __bfloat162float(q[1].h[j]) + __bfloat162float(q[2].h[j]) +
                       __bfloat162float(q[3].h[j]));
    return;
  }
  float t[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) t[j] = v[j];
  shortcut_add8_slow(res, Cres, RH, RW, res_pool, F, b, oy, ox, c0, t);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = t[j];
}
#endif

}  // namespace fn
