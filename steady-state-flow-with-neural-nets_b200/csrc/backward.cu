// Backward pieces of the flow_net training step (reference: flow_net.train -> AdamOptimizer.minimize,
// model/flow_net.py:44-46, i.e. TF autodiff of flow_architecture.py:129-248).  Data gradients of the
// convolutions reuse the FORWARD kernels (a stride-1 conv's dgrad is a stride-1 conv with flipped,
// transposed weights; a stride-2 conv's dgrad is the transposed conv with the same weights and vice
// versa), so what lives here is:
//   * the pointwise adjoints of the fused prologues / epilogues (concat_elu & co, gate, shortcut,
//     tanh + l2 loss),
//   * the weight / bias gradient reductions (CUDA-core, correctness-first: two-stage, deterministic).
// Activation gradients are stored bf16 (like activations), parameter gradients fp32.
#include "common.cuh"

namespace fn {

constexpr int BW_THREADS = 256;
static inline int bw_grid(long long items) {
  long long g = (items + BW_THREADS - 1) / BW_THREADS;
  const long long cap = 148LL * 16;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

// ---- gate: y = u * sigmoid(v)  ->  du = g*s, dv = g*u*s*(1-s)          (flow_architecture.py:150-151)
__global__ void k_gate_bwd(const __nv_bfloat16* __restrict__ c2, const __nv_bfloat16* __restrict__ g,
                           __nv_bfloat16* __restrict__ dc2, long long npix, int F) {
  const long long total = npix * F;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / F;
    const int c = (int)(i % F);
    const float u = __bfloat162float(c2[p * 2 * F + c]);
    const float v = __bfloat162float(c2[p * 2 * F + F + c]);
    const float gg = __bfloat162float(g[i]);
    const float s = 1.f / (1.f + __expf(-v));
    dc2[p * 2 * F + c] = __float2bfloat16(gg * s);
    dc2[p * 2 * F + F + c] = __float2bfloat16(gg * u * s * (1.f - s));
  }
}

// ---- prologue adjoint: dx (+)= dA0 * f'(x) [- dA1 * f'(-x)]             (flow_architecture.py:14-29, 143-145)
__device__ __forceinline__ float elu_grad(float z) { return z > 0.f ? 1.f : __expf(z); }
__global__ void k_prologue_bwd(const __nv_bfloat16* __restrict__ dA, const __nv_bfloat16* __restrict__ x,
                               __nv_bfloat16* __restrict__ dx, long long npix, int C, int kind, int accumulate,
                               float keep_prob, unsigned long long seed, const unsigned long long* seed_dev) {
  if (seed_dev) seed += *seed_dev;
  const int mult = pro_mult(kind);
  const long long total = npix * C;
  const bool drop = keep_prob < 1.0f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / C;
    const int c = (int)(i % C);
    const float xv = __bfloat162float(x[i]);
    float a0 = __bfloat162float(dA[p * mult * C + c]);
    float a1 = mult == 2 ? __bfloat162float(dA[p * mult * C + C + c]) : 0.f;
    if (drop) {   // the forward scaled prologue output k of pixel p by dropout_scale(seed, p*Keff + k)
      a0 *= dropout_scale(seed, (unsigned long long)(p * mult * C + c), keep_prob);
      if (mult == 2) a1 *= dropout_scale(seed, (unsigned long long)(p * mult * C + C + c), keep_prob);
    }
    float d;
    switch (kind) {
      case FN_PRO_CONCAT_ELU:  d = a0 * elu_grad(xv) - a1 * elu_grad(-xv); break;
      case FN_PRO_ELU:         d = a0 * elu_grad(xv); break;
      case FN_PRO_CONCAT_RELU: d = (xv > 0.f ? a0 : 0.f) - (xv < 0.f ? a1 : 0.f); break;
      case FN_PRO_RELU:        d = xv > 0.f ? a0 : 0.f; break;
      default:                 d = a0; break;
    }
    if (accumulate) d += __bfloat162float(dx[i]);
    dx[i] = __float2bfloat16(d);
  }
}

// ---- shortcut adjoint: out = pad_front(avgpool?(x)) + ...  ->  dx (+)= slice(g) [/4 broadcast]   (:153-165)
__global__ void k_shortcut_bwd(const __nv_bfloat16* __restrict__ g, __nv_bfloat16* __restrict__ dx, int B, int OH, int OW,
                               int F, int H, int W, int Cin, int pooled, int accumulate) {
  const long long total = (long long)B * H * W * Cin;
  const int shift = F - Cin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cin);
    long long r = i / Cin;
    const int xw = (int)(r % W);
    r /= W;
    const int yh = (int)(r % H);
    const int b = (int)(r / H);
    float d;
    if (pooled) {
      const int oy = yh >> 1, ox = xw >> 1;
      const int cnt = (min(2 * oy + 2, H) - 2 * oy) * (min(2 * ox + 2, W) - 2 * ox);   // TF 'SAME' edge windows
      d = __bfloat162float(g[(((long long)b * OH + oy) * OW + ox) * F + shift + c]) / (float)cnt;
    } else {
      d = __bfloat162float(g[(((long long)b * OH + yh) * OW + xw) * F + shift + c]);
    }
    if (accumulate) d += __bfloat162float(dx[i]);
    dx[i] = __float2bfloat16(d);
  }
}


// ---- 8-channel (16-byte) variants of the three element-wise adjoints above, used whenever the channel counts allow
__device__ __forceinline__ void ld8(bf16x8& r, const __nv_bfloat16* p) { r.v = *reinterpret_cast<const uint4*>(p); }
__global__ void __launch_bounds__(BW_THREADS) k_gate_bwd_v8(const __nv_bfloat16* __restrict__ c2, const __nv_bfloat16* __restrict__ g,
                                                            __nv_bfloat16* __restrict__ dc2, long long npix, int F) {
  const int fc = F >> 3;
  const long long total = npix * fc;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / fc;
    const int c = (int)(i - p * fc) * 8;
    bf16x8 u, v, gg, du, dv;
    ld8(u, c2 + p * 2 * F + c);
    ld8(v, c2 + p * 2 * F + F + c);
    ld8(gg, g + p * F + c);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float sg = 1.f / (1.f + __expf(-__bfloat162float(v.h[j])));
      const float gs = __bfloat162float(gg.h[j]) * sg;
      du.h[j] = __float2bfloat16(gs);
      dv.h[j] = __float2bfloat16(gs * __bfloat162float(u.h[j]) * (1.f - sg));
    }
    *reinterpret_cast<uint4*>(dc2 + p * 2 * F + c) = du.v;
    *reinterpret_cast<uint4*>(dc2 + p * 2 * F + F + c) = dv.v;
  }
}

template <bool DROP>
__global__ void __launch_bounds__(BW_THREADS) k_prologue_bwd_v8(const __nv_bfloat16* __restrict__ dA, const __nv_bfloat16* __restrict__ x,
                                                                __nv_bfloat16* __restrict__ dx, long long npix, int C, int kind,
                                                                int accumulate, float keep_prob, unsigned long long seed,
                                                                const unsigned long long* seed_dev) {
  if (DROP && seed_dev) seed += *seed_dev;
  const int mult = pro_mult(kind);
  const int cc = C >> 3;
  const long long total = npix * cc;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / cc;
    const int c = (int)(i - p * cc) * 8;
    bf16x8 xv, a0, a1, old, o;
    ld8(xv, x + p * C + c);
    ld8(a0, dA + p * mult * C + c);
    if (mult == 2) ld8(a1, dA + p * mult * C + C + c);
    if (accumulate) ld8(old, dx + p * C + c);
    float sc0[8], sc1[8];
    if (DROP) {
      dropout_scale8(seed, (unsigned long long)(p * mult * C + c), keep_prob, sc0);
      if (mult == 2) dropout_scale8(seed, (unsigned long long)(p * mult * C + C + c), keep_prob, sc1);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xf = __bfloat162float(xv.h[j]);
      float f0 = __bfloat162float(a0.h[j]);
      float f1 = mult == 2 ? __bfloat162float(a1.h[j]) : 0.f;
      if (DROP) {
        f0 *= sc0[j];
        if (mult == 2) f1 *= sc1[j];
      }
      float d;
      switch (kind) {
        case FN_PRO_CONCAT_ELU: {   // elu'(x) = x>0 ? 1 : e^x and elu'(-x) = x<0 ? 1 : e^-x share one exponential
          const float e = __expf(-fabsf(xf));
          d = xf > 0.f ? f0 - f1 * e : f0 * e - f1;
          break;
        }
        case FN_PRO_ELU:         d = f0 * elu_grad(xf); break;
        case FN_PRO_CONCAT_RELU: d = (xf > 0.f ? f0 : 0.f) - (xf < 0.f ? f1 : 0.f); break;
        case FN_PRO_RELU:        d = xf > 0.f ? f0 : 0.f; break;
        default:                 d = f0; break;
      }
      if (accumulate) d += __bfloat162float(old.h[j]);
      o.h[j] = __float2bfloat16(d);
    }
    *reinterpret_cast<uint4*>(dx + p * C + c) = o.v;
  }
}

__global__ void __launch_bounds__(BW_THREADS) k_shortcut_bwd_v8(const __nv_bfloat16* __restrict__ g, __nv_bfloat16* __restrict__ dx, int B,
                                                                int OH, int OW, int F, int H, int W, int Cin, int pooled,
                                                                int accumulate) {
  const int cc = Cin >> 3;
  const long long total = (long long)B * H * W * cc;
  const int shift = F - Cin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i / cc;
    const int c = (int)(i - r * cc) * 8;
    const long long pix = r;
    const int xw = (int)(r % W);
    r /= W;
    const int yh = (int)(r % H);
    const int b = (int)(r / H);
    float scale = 1.f;
    int oy = yh, ox = xw;
    if (pooled) {
      oy = yh >> 1; ox = xw >> 1;
      scale = 1.f / (float)((min(2 * oy + 2, H) - 2 * oy) * (min(2 * ox + 2, W) - 2 * ox));   // TF 'SAME' edge windows
    }
    bf16x8 gv, old, o;
    ld8(gv, g + (((long long)b * OH + oy) * OW + ox) * F + shift + c);
    if (accumulate) ld8(old, dx + pix * Cin + c);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float d = __bfloat162float(gv.h[j]) * scale;
      if (accumulate) d += __bfloat162float(old.h[j]);
      o.h[j] = __float2bfloat16(d);
    }
    *reinterpret_cast<uint4*>(dx + pix * Cin + c) = o.v;
  }
}

// bias gradient, 8 channels per thread: blockDim = (N/8) x rows
__global__ void __launch_bounds__(BW_THREADS) k_bias_grad_v8(const __nv_bfloat16* __restrict__ dY, float* __restrict__ partial,
                                                             long long npix, int N) {
  __shared__ float sm[BW_THREADS][9];
  const int nc = N >> 3;
  const int per = BW_THREADS / nc;
  const int c = threadIdx.x % nc, rg = threadIdx.x / nc;
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (rg < per) {
    const long long step = (long long)gridDim.x * per;
    long long pp = (long long)blockIdx.x * per + rg;
    for (; pp + 3 * step < npix; pp += 4 * step) {      // four independent 16-byte loads in flight
      bf16x8 v0, v1, v2, v3;
      ld8(v0, dY + pp * N + c * 8);
      ld8(v1, dY + (pp + step) * N + c * 8);
      ld8(v2, dY + (pp + 2 * step) * N + c * 8);
      ld8(v3, dY + (pp + 3 * step) * N + c * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        s[j] += (__bfloat162float(v0.h[j]) + __bfloat162float(v1.h[j])) + (__bfloat162float(v2.h[j]) + __bfloat162float(v3.h[j]));
    }
    for (; pp < npix; pp += step) {
      bf16x8 v;
      ld8(v, dY + pp * N + c * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += __bfloat162float(v.h[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sm[threadIdx.x][j] = s[j];
  __syncthreads();
  if (threadIdx.x < N) {
    const int cc = threadIdx.x >> 3, j = threadIdx.x & 7;
    float a = 0.f;
    for (int r = 0; r < per; ++r) a += sm[r * nc + cc][j];
    partial[(long long)blockIdx.x * N + threadIdx.x] = a;
  }
}


// ---- data gradient of the nin skip (x_1 += nin(pro(a)), flow_architecture.py:136-142) fused with the adjoint of its
// prologue: dA[k] = sum_f dy[f] * Wn[k][f] (k < 2 Ca), d(a)[c] (+)= dA[c] f'(a[c]) - dA[Ca + c] f'(-a[c]).
// A 1x1 conv with at most 64 x 128 weights: CUDA cores; a warp = 32 pixels of one 8-channel chunk, so the fp32 weights
// (staged in shared memory) are broadcast reads and every thread keeps its pixel's dy row in registers.
template <int F>
__global__ void __launch_bounds__(256) k_nin_bwd(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ w,
                                                 const __nv_bfloat16* __restrict__ a, __nv_bfloat16* __restrict__ da, long long npix,
                                                 int Ca, int kind, int accumulate) {
  extern __shared__ __align__(16) float ws[];         // [2 Ca][F]
  for (int i = threadIdx.x; i < 2 * Ca * F; i += 256) ws[i] = w[i];
  __syncthreads();
  const int chunks = Ca >> 3;
  const long long total = npix * chunks;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int chunk = (int)(i / npix);
    const long long pix = i - (long long)chunk * npix;
    float g[F];
#pragma unroll
    for (int f8 = 0; f8 < F; f8 += 8) {
      bf16x8 v;
      ld8(v, dy + pix * F + f8);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[f8 + j] = __bfloat162float(v.h[j]);
    }
    bf16x8 av, old, out;
    ld8(av, a + pix * Ca + chunk * 8);
    if (accumulate) ld8(old, da + pix * Ca + chunk * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4* w0 = reinterpret_cast<const float4*>(ws + (chunk * 8 + j) * F);       // rows are 32-byte aligned
      const float4* w1 = reinterpret_cast<const float4*>(ws + (Ca + chunk * 8 + j) * F);
      float d0 = 0.f, d1 = 0.f;
#pragma unroll
      for (int f4 = 0; f4 < F / 4; ++f4) {
        const float4 u = w0[f4], v = w1[f4];
        d0 = fmaf(g[4 * f4], u.x, d0); d0 = fmaf(g[4 * f4 + 1], u.y, d0); d0 = fmaf(g[4 * f4 + 2], u.z, d0); d0 = fmaf(g[4 * f4 + 3], u.w, d0);
        d1 = fmaf(g[4 * f4], v.x, d1); d1 = fmaf(g[4 * f4 + 1], v.y, d1); d1 = fmaf(g[4 * f4 + 2], v.z, d1); d1 = fmaf(g[4 * f4 + 3], v.w, d1);
      }
      d0 = __bfloat162float(__float2bfloat16(d0));    // the two-launch path stores dA as bf16
      d1 = __bfloat162float(__float2bfloat16(d1));
      const float xf = __bfloat162float(av.h[j]);
      float d;
      if (kind == FN_PRO_CONCAT_ELU) {
        const float e = __expf(-fabsf(xf));
        d = xf > 0.f ? d0 - d1 * e : d0 * e - d1;
      } else {                                        // FN_PRO_CONCAT_RELU
        d = (xf > 0.f ? d0 : 0.f) - (xf < 0.f ? d1 : 0.f);
      }
      if (accumulate) d += __bfloat162float(old.h[j]);
      out.h[j] = __float2bfloat16(d);
    }
    *reinterpret_cast<uint4*>(da + pix * Ca + chunk * 8) = out.v;
  }
}

template <int F>
static int nin_bwd_launch(const void* dy, const float* w, const void* a, void* da, long long npix, int Ca, int kind, int accumulate,
                          cudaStream_t s) {
  const size_t smem = (size_t)2 * Ca * F * sizeof(float);
  static bool configured = false;
  if (!configured && smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_nin_bwd<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(nin_bwd): %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  k_nin_bwd<F><<<bw_grid(npix * (Ca / 8)), 256, smem, s>>>((const __nv_bfloat16*)dy, w, (const __nv_bfloat16*)a, (__nv_bfloat16*)da, npix,
                                                            Ca, kind, accumulate);
  count_launch();
  return check_launch("k_nin_bwd");
}

// ---- tanh + l2 loss: loss += sum((p-t)^2)/2 ; dpre[pix, 0..1] = (p-t)*(1-p^2), channels 2..7 = 0
__global__ void k_tanh_l2_bwd(const float* __restrict__ p, const float* __restrict__ t, float* loss,
                              __nv_bfloat16* __restrict__ dpre, long long npix) {
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
    bf16x8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) o.h[j] = __float2bfloat16(0.f);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const float pv = p[i * 2 + c];
      const float d = pv - t[i * 2 + c];
      s = fmaf(d, d, s);
      o.h[c] = __float2bfloat16(d * (1.f - pv * pv));
    }
    *reinterpret_cast<uint4*>(dpre + i * 8) = o.v;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ float ws[BW_THREADS / 32];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < BW_THREADS / 32 ? ws[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0 && loss) atomicAdd(loss, 0.5f * s);
  }
}

// ---- deterministic second stage: out[i] = sum_s partial[s][i], for one or two (partial, out) pairs in one launch (a
// conv's weight and bias gradients).  block = 32 outputs x TY split lanes: lane y adds partials y, y+TY, .. with four
// independent loads in flight, then the TY lane sums are added in a fixed order -- the result depends on S and TY
// only, not on the launch geometry of stage 1.  TY = 32 for small jobs (a lone block walking 592 partials eight lanes
// wide took 15 us: 18 dependent round trips to memory), TY = 8 for the large ones.
struct ReduceJob {
  const float* partial;
  float* out;
  long long n;
  int S;
};
template <int TY>
__global__ void __launch_bounds__(32 * TY) k_reduce_partials(ReduceJob j0, ReduceJob j1, unsigned nb0) {
  __shared__ float sm[TY][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const bool second = blockIdx.x >= nb0;
  const float* __restrict__ partial = second ? j1.partial : j0.partial;
  float* __restrict__ out = second ? j1.out : j0.out;
  const long long n = second ? j1.n : j0.n;
  const int S = second ? j1.S : j0.S;
  const long long i = (long long)(blockIdx.x - (second ? nb0 : 0u)) * 32 + tx;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (i < n) {
    int k = ty;
    for (; k + 3 * TY < S; k += 4 * TY) {
      a0 += partial[(long long)k * n + i];
      a1 += partial[(long long)(k + TY) * n + i];
      a2 += partial[(long long)(k + 2 * TY) * n + i];
      a3 += partial[(long long)(k + 3 * TY) * n + i];
    }
    for (; k < S; k += TY) a0 += partial[(long long)k * n + i];
  }
  sm[ty][tx] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (ty == 0 && i < n) {
    float s = 0.f;
#pragma unroll
    for (int y = 0; y < TY; ++y) s += sm[y][tx];
    out[i] = s;
  }
}

// ---- bias gradient: db[n] = sum_p dY[p][n]; stage 1 writes partial[blockIdx.x][n]
__global__ void k_bias_grad(const __nv_bfloat16* __restrict__ dY, float* __restrict__ partial, long long npix, int N) {
  // thread -> column n = tid % N (N <= 256 divides 256 for the network's widths, else tail threads idle)
  const int per = BW_THREADS / N > 0 ? BW_THREADS / N : 1;
  const int n = threadIdx.x % N, rg = threadIdx.x / N;
  float s = 0.f;
  if (rg < per && threadIdx.x < per * N) {
    const long long p0 = (long long)blockIdx.x * per + rg;
    for (long long pp = p0; pp < npix; pp += (long long)gridDim.x * per) s += __bfloat162float(dY[pp * N + n]);
  }
  __shared__ float sm[BW_THREADS];
  sm[threadIdx.x] = (rg < per && threadIdx.x < per * N) ? s : 0.f;
  __syncthreads();
  if (threadIdx.x < N) {
    float a = 0.f;
    for (int r = 0; r < per; ++r) a += sm[r * N + threadIdx.x];
    partial[(long long)blockIdx.x * N + threadIdx.x] = a;
  }
}

// ---- weight gradient of a 'SAME' conv: dW[tap][k][n] = sum_pix pro(x)[pix*s + tap - pad][k] * dY[pix][n]
// CTA = one (tap, 16 k, 16 n) tile over one slice of the pixels; thread = one (k, n) pair.
struct WgradParams {
  const __nv_bfloat16* x;
  const __nv_bfloat16* dY;
  float* partial;      // [S][T][Keff][N]
  int B, H, W, C, OH, OW, N, ksize, stride, pad_t, pad_l, pro, Keff;
  int transposed;      // 1: transposed conv: x is the INPUT [B,H,W,C] and dY the 2x output [B,OH,OW,N]
  float keep_prob;
  unsigned long long seed;
  const unsigned long long* seed_dev;
  int kt, nt;          // number of 16-wide k / n tiles
};

constexpr int WG_PIX = 32;

__global__ void __launch_bounds__(256) k_conv_wgrad(const WgradParams p) {
  __shared__ float sa[WG_PIX][17];
  __shared__ float sd[WG_PIX][17];
  int tile = blockIdx.x;
  const int nti = tile % p.nt;
  tile /= p.nt;
  const int kti = tile % p.kt;
  const int tap = tile / p.kt;
  const int di = tap / p.ksize, dj = tap % p.ksize;
  const int kl = threadIdx.x >> 4, nl = threadIdx.x & 15;     // (k, n) owned by this thread
  const int mult = pro_mult(p.pro);
  // GEMM-row pixels: conv -> output pixels (OH x OW); transposed conv -> input pixels (H x W)
  const int RH = p.transposed ? p.H : p.OH, RW = p.transposed ? p.W : p.OW;
  const long long rows = (long long)p.B * RH * RW;
  const long long per_split = (rows + gridDim.y - 1) / gridDim.y;
  const long long r0 = (long long)blockIdx.y * per_split;
  const long long r1 = min(rows, r0 + per_split);
  float acc = 0.f;
  for (long long base = r0; base < r1; base += WG_PIX) {
    // stage WG_PIX rows: 16 activations (k tile) and 16 output-gradient values (n tile) each
    for (int i = threadIdx.x; i < WG_PIX * 32; i += 256) {
      const int pr = i >> 5, col = i & 31;
      const long long row = base + pr;
      float v = 0.f;
      if (row < r1) {
        const int rx = (int)(row % RW);
        const int ry = (int)((row / RW) % RH);
        const int b = (int)(row / ((long long)RW * RH));
        if (col < 16) {                       // activation for k = kti*16 + col
          const int k = kti * 16 + col;
          if (k < p.Keff) {
            int iy, ix;
            if (p.transposed) { iy = ry; ix = rx; }
            else { iy = ry * p.stride + di - p.pad_t; ix = rx * p.stride + dj - p.pad_l; }
            if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) {
              const int c = k >= p.C ? k - p.C : k;
              const long long pidx = ((long long)b * p.H + iy) * p.W + ix;
              float p0, p1;
              pro_apply(p.pro, __bfloat162float(p.x[pidx * p.C + c]), p0, p1);
              v = (mult == 2 && k >= p.C) ? p1 : p0;
              if (p.pro != FN_PRO_NONE) v = __bfloat162float(__float2bfloat16(v));   // the forward operand was bf16
              if (p.keep_prob < 1.0f)
                v *= dropout_scale(p.seed + (p.seed_dev ? *p.seed_dev : 0ull), (unsigned long long)(pidx * p.Keff + k), p.keep_prob);
            }
          }
          sa[pr][col] = v;
        } else {                              // output gradient for n = nti*16 + (col-16)
          const int n = nti * 16 + (col - 16);
          if (n < p.N) {
            int oy, ox;
            bool ok = true;
            if (p.transposed) {               // out[2i+di, 2j+dj] <- in[i,j] * W[di,dj]   
              oy = 2 * ry + di; ox = 2 * rx + dj;
              ok = oy < p.OH && ox < p.OW;
            } else { oy = ry; ox = rx; }
            if (ok) v = __bfloat162float(p.dY[(((long long)b * p.OH + oy) * p.OW + ox) * p.N + n]);
          }
          sd[pr][col - 16] = v;
        }
      } else {
        if (col < 16) sa[pr][col] = 0.f; else sd[pr][col - 16] = 0.f;
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int pr = 0; pr < WG_PIX; ++pr) acc = fmaf(sa[pr][kl], sd[pr][nl], acc);
    __syncthreads();
  }
  const int k = kti * 16 + kl, n = nti * 16 + nl;
  if (k < p.Keff && n < p.N) {
    const long long T = (long long)p.ksize * p.ksize;
    if (p.transposed) p.partial[(((long long)blockIdx.y * T + tap) * p.N + n) * p.Keff + k] = acc;   // TF layout [k,k,Cout,Cin]
    else p.partial[(((long long)blockIdx.y * T + tap) * p.Keff + k) * p.N + n] = acc;
  }
}

// ---- weight gradient of the network's first conv (Cin = 1, no prologue, 3x3 stride 1; flow_architecture.py:177 ->
// res_block conv_1 on the boundary): dW[tap][n] = sum_pix x[pix + tap - 1] * dY[pix][n].  One thread walks pixels with
// 9 x 8 accumulators (blockIdx.y = 8-channel chunk of n); block tree reduction, one partial per block.
constexpr int WG1_BLOCKS = 296;
__global__ void __launch_bounds__(256) k_wgrad_cin1(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dY,
                                                    float* __restrict__ partial, int B, int H, int W, int N) {
  __shared__ float red[8][72];
  const int n0 = blockIdx.y * 8;
  const long long npix = (long long)B * H * W;
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += (long long)gridDim.x * blockDim.x) {
    const int px = (int)(pix % W);
    const int py = (int)((pix / W) % H);
    bf16x8 g;
    g.v = *reinterpret_cast<const uint4*>(dY + pix * N + n0);
    float gv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) gv[j] = __bfloat162float(g.h[j]);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int iy = py + t / 3 - 1, ix = px + t % 3 - 1;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
        const float xv = __bfloat162float(x[pix + (long long)(t / 3 - 1) * W + (t % 3 - 1)]);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[t][j] = fmaf(xv, gv[j], acc[t][j]);
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = acc[t][j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red[warp][t * 8 + j] = v;
    }
  __syncthreads();
  if (threadIdx.x < 72) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
    const int t = threadIdx.x >> 3, j = threadIdx.x & 7;
    partial[((long long)blockIdx.x * 9 + t) * N + n0 + j] = v;
  }
}

static int reduce_partials(const float* partial, float* out, long long n, int S, cudaStream_t s, const float* partial2 = nullptr,
                           float* out2 = nullptr, long long n2 = 0, int S2 = 0) {
  const ReduceJob j0{partial, out, n, S}, j1{partial2, out2, partial2 ? n2 : 0, S2};
  const unsigned nb0 = (unsigned)((n + 31) / 32), nb1 = (unsigned)((j1.n + 31) / 32);
  if (nb0 + nb1 <= 592 && (S > 32 || S2 > 32))
    k_reduce_partials<32><<<nb0 + nb1, 1024, 0, s>>>(j0, j1, nb0);
  else
    k_reduce_partials<8><<<nb0 + nb1, 256, 0, s>>>(j0, j1, nb0);
  count_launch();
  return check_launch("k_reduce_partials");
}

int gate_bwd(const void* c2, const void* g, void* dc2, long long npix, int F, cudaStream_t s) {
  if (npix == 0) return 0;
  if ((F & 7) == 0)
    k_gate_bwd_v8<<<bw_grid(npix * (F / 8)), BW_THREADS, 0, s>>>((const __nv_bfloat16*)c2, (const __nv_bfloat16*)g, (__nv_bfloat16*)dc2,
                                                                   npix, F);
  else
    k_gate_bwd<<<bw_grid(npix * F), BW_THREADS, 0, s>>>((const __nv_bfloat16*)c2, (const __nv_bfloat16*)g, (__nv_bfloat16*)dc2, npix, F);
  count_launch();
  return check_launch("k_gate_bwd");
}

int prologue_bwd(const void* dA, const void* x, void* dx, long long npix, int C, int kind, int accumulate, float keep_prob,
                 unsigned long long seed, const unsigned long long* seed_dev, cudaStream_t s) {
  if (npix == 0) return 0;
  if ((C & 7) == 0)
    if (keep_prob < 1.0f)
      k_prologue_bwd_v8<true><<<bw_grid(npix * (C / 8)), BW_THREADS, 0, s>>>((const __nv_bfloat16*)dA, (const __nv_bfloat16*)x,
                                                                               (__nv_bfloat16*)dx, npix, C, kind, accumulate, keep_prob, seed, seed_dev);
    else
      k_prologue_bwd_v8<false><<<bw_grid(npix * (C / 8)), BW_THREADS, 0, s>>>((const __nv_bfloat16*)dA, (const __nv_bfloat16*)x,
                                                                                (__nv_bfloat16*)dx, npix, C, kind, accumulate, keep_prob, seed, seed_dev);
  else
    k_prologue_bwd<<<bw_grid(npix * C), BW_THREADS, 0, s>>>((const __nv_bfloat16*)dA, (const __nv_bfloat16*)x, (__nv_bfloat16*)dx,
                                                            npix, C, kind, accumulate, keep_prob, seed, seed_dev);
  count_launch();
  return check_launch("k_prologue_bwd");
}

int shortcut_bwd(const void* g, void* dx, int B, int OH, int OW, int F, int H, int W, int Cin, int pooled, int accumulate,
                 cudaStream_t s) {
  const long long total = (long long)B * H * W * Cin;
  if (total == 0) return 0;
  if ((Cin & 7) == 0 && (F & 7) == 0)
    k_shortcut_bwd_v8<<<bw_grid(total / 8), BW_THREADS, 0, s>>>((const __nv_bfloat16*)g, (__nv_bfloat16*)dx, B, OH, OW, F, H, W, Cin,
                                                                  pooled, accumulate);
  else
    k_shortcut_bwd<<<bw_grid(total), BW_THREADS, 0, s>>>((const __nv_bfloat16*)g, (__nv_bfloat16*)dx, B, OH, OW, F, H, W, Cin,
                                                         pooled, accumulate);
  count_launch();
  return check_launch("k_shortcut_bwd");
}

int tanh_l2_bwd(const float* p, const float* t, float* loss, void* dpre, long long npix, cudaStream_t s) {
  if (npix == 0) return 0;
  k_tanh_l2_bwd<<<bw_grid(npix), BW_THREADS, 0, s>>>(p, t, loss, (__nv_bfloat16*)dpre, npix);
  count_launch();
  return check_launch("k_tanh_l2_bwd");
}

long long bias_grad_workspace(long long npix, int N) {
  (void)npix;
  return (long long)148 * 4 * N;
}
int bias_grad(const void* dY, float* db, float* workspace, long long npix, int N, cudaStream_t s) {
  FN_REQUIRE(N >= 1 && N <= BW_THREADS, FN_E_UNSUPPORTED, "bias_grad: N=%d out of range", N);
  const int S = 148 * 4;
  if ((N & 7) == 0) k_bias_grad_v8<<<S, BW_THREADS, 0, s>>>((const __nv_bfloat16*)dY, workspace, npix, N);
  else k_bias_grad<<<S, BW_THREADS, 0, s>>>((const __nv_bfloat16*)dY, workspace, npix, N);
  count_launch();
  int rc = check_launch("k_bias_grad");
  if (rc) return rc;
  return reduce_partials(workspace, db, N, S, s);
}

static int wgrad_splits(long long rows) {
  long long S = rows / 2048;
  if (S < 1) S = 1;
  if (S > 64) S = 64;
  return (int)S;
}
long long wgrad_workspace(long long rows, int T, int Keff, int N) { return (long long)wgrad_splits(rows) * T * Keff * N; }

// wgrad_tc.cu
long long wgrad_tc_workspace(int B, int H, int W, int C, int OH, int OW, int N, int ksize, int stride, int pro);
bool wgrad_tc_try(const void* x, const void* dY, float* partial, int B, int H, int W, int C, int OH, int OW, int N, int ksize,
                  int stride, int pro, float keep_prob, unsigned long long seed, int* S_out, cudaStream_t stream, int& rc,
                  bool want_bias = false, float** bias_partial_out = nullptr, const unsigned long long* seed_dev = nullptr);
static long long max_ll(long long a, long long b) { return a > b ? a : b; }

int conv_wgrad_launch(WgradParams& p, float* dW, float* workspace, cudaStream_t s) {
  const int T = p.ksize * p.ksize;
  p.kt = (p.Keff + 15) / 16;
  p.nt = (p.N + 15) / 16;
  const long long rows = (long long)p.B * (p.transposed ? (long long)p.H * p.W : (long long)p.OH * p.OW);
  const int S = wgrad_splits(rows);
  p.partial = workspace;
  dim3 grid((unsigned)(T * p.kt * p.nt), (unsigned)S);
  k_conv_wgrad<<<grid, 256, 0, s>>>(p);
  count_launch();
  int rc = check_launch("k_conv_wgrad");
  if (rc) return rc;
  return reduce_partials(workspace, dW, (long long)T * p.Keff * p.N, S, s);
}

}  // namespace fn

using namespace fn;

extern "C" {

int fn_gate_bwd(const void* c2, const void* g, void* dc2, int64_t npix, int32_t F, void* stream) {
  FN_REQUIRE(c2 && g && dc2 && npix >= 0 && F > 0, FN_E_INVAL, "fn_gate_bwd: bad arguments");
  return gate_bwd(c2, g, dc2, npix, F, (cudaStream_t)stream);
}

int fn_prologue_bwd(const void* dA, const void* x, void* dx, int64_t npix, int32_t C, int32_t kind, int32_t accumulate,
                    float keep_prob, uint64_t seed, void* stream) {
  FN_REQUIRE(dA && x && dx && npix >= 0 && C > 0, FN_E_INVAL, "fn_prologue_bwd: bad arguments");
  FN_REQUIRE(kind >= FN_PRO_NONE && kind <= FN_PRO_RELU, FN_E_INVAL, "fn_prologue_bwd: bad kind %d", kind);
  return prologue_bwd(dA, x, dx, npix, C, kind, accumulate, keep_prob, seed, nullptr, (cudaStream_t)stream);
}

int fn_prologue_bwd_dev(const void* dA, const void* x, void* dx, int64_t npix, int32_t C, int32_t kind, int32_t accumulate,
                        float keep_prob, uint64_t seed, const uint64_t* seed_dev, void* stream) {
  FN_REQUIRE(dA && x && dx && npix >= 0 && C > 0, FN_E_INVAL, "fn_prologue_bwd_dev: bad arguments");
  FN_REQUIRE(kind >= FN_PRO_NONE && kind <= FN_PRO_RELU, FN_E_INVAL, "fn_prologue_bwd_dev: bad kind %d", kind);
  return prologue_bwd(dA, x, dx, npix, C, kind, accumulate, keep_prob, seed, (const unsigned long long*)seed_dev, (cudaStream_t)stream);
}

/* data gradient of the nin skip fused with the adjoint of its prologue (concat kinds): dy bf16 [npix,F], w fp32 [2 Ca, F]
 * (the variable itself, TF [dim, hiddens]), a bf16 [npix,Ca] -> da bf16 [npix,Ca] (+= if accumulate) */
int fn_nin_bwd(const void* dy, const float* w, const void* a, void* da, int64_t npix, int32_t F, int32_t Ca, int32_t kind,
               int32_t accumulate, void* stream) {
  FN_REQUIRE(dy && w && a && da && npix >= 0, FN_E_INVAL, "fn_nin_bwd: bad arguments");
  FN_REQUIRE(kind == FN_PRO_CONCAT_ELU || kind == FN_PRO_CONCAT_RELU, FN_E_UNSUPPORTED, "fn_nin_bwd: concat prologues only (kind %d)", kind);
  FN_REQUIRE((Ca & 7) == 0 && Ca >= 8 && (size_t)2 * Ca * F * 4 <= 64 * 1024, FN_E_UNSUPPORTED, "fn_nin_bwd: Ca=%d F=%d", Ca, F);
  if (npix == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  switch (F) {
    case 8: return nin_bwd_launch<8>(dy, w, a, da, npix, Ca, kind, accumulate, s);
    case 16: return nin_bwd_launch<16>(dy, w, a, da, npix, Ca, kind, accumulate, s);
    case 32: return nin_bwd_launch<32>(dy, w, a, da, npix, Ca, kind, accumulate, s);
    case 64: return nin_bwd_launch<64>(dy, w, a, da, npix, Ca, kind, accumulate, s);
    default: set_error("fn_nin_bwd: F=%d (8, 16, 32, 64)", F); return FN_E_UNSUPPORTED;
  }
}

int fn_shortcut_bwd(const void* g, void* dx, int32_t B, int32_t OH, int32_t OW, int32_t F, int32_t H, int32_t W, int32_t Cin,
                    int32_t pooled, int32_t accumulate, void* stream) {
  FN_REQUIRE(g && dx && B > 0 && OH > 0 && OW > 0 && F > 0 && H > 0 && W > 0 && Cin > 0 && Cin <= F, FN_E_INVAL,
             "fn_shortcut_bwd: bad arguments");
  return shortcut_bwd(g, dx, B, OH, OW, F, H, W, Cin, pooled, accumulate, (cudaStream_t)stream);
}

int fn_tanh_l2_bwd(const float* p, const float* t, float* loss, void* dpre, int64_t npix, void* stream) {
  FN_REQUIRE(p && t && dpre && npix >= 0, FN_E_INVAL, "fn_tanh_l2_bwd: bad arguments");
  return tanh_l2_bwd(p, t, loss, dpre, npix, (cudaStream_t)stream);
}

int64_t fn_bias_grad_workspace(int64_t npix, int32_t N) { return bias_grad_workspace(npix, N); }
int fn_bias_grad(const void* dY, float* db, float* workspace, int64_t npix, int32_t N, void* stream) {
  FN_REQUIRE(dY && db && workspace && npix >= 0, FN_E_INVAL, "fn_bias_grad: bad arguments");
  return bias_grad(dY, db, workspace, npix, N, (cudaStream_t)stream);
}

int64_t fn_conv_wgrad_workspace(const fn_conv_desc* d) {
  if (!d) return 0;
  const int OH = (d->H + d->stride - 1) / d->stride, OW = (d->W + d->stride - 1) / d->stride;
  return max_ll(max_ll(max_ll(wgrad_workspace((long long)d->B * OH * OW, d->ksize * d->ksize, d->Cin * pro_mult(d->prologue), d->Cout),
                              (long long)WG1_BLOCKS * 9 * d->Cout),
                       bias_grad_workspace((long long)d->B * OH * OW, d->Cout)),
                wgrad_tc_workspace(d->B, d->H, d->W, d->Cin, OH, OW, d->Cout, d->ksize, d->stride, d->prologue));
}

/* d describes the FORWARD conv (x, dims, ksize, stride, prologue, keep_prob/seed; w, bias, y are ignored);
 * dY: bf16 [B,OH,OW,Cout]; dW: fp32 [k*k, Keff, Cout] (TF HWIO flattened) */
static int conv_wgrad_impl(const fn_conv_desc* d, const void* dY, float* dW, float* db, float* workspace, void* stream);

int fn_conv_wgrad(const fn_conv_desc* d, const void* dY, float* dW, float* workspace, void* stream) {
  return conv_wgrad_impl(d, dY, dW, nullptr, workspace, stream);
}

/* the same plus the bias gradient db[n] = sum_pix dY[pix][n] (fp32 [Cout]): on the tensor-core path it is one more
 * accumulator of the same kernel (a plane of ones as the A operand), otherwise the bias kernel runs after it */
int fn_conv_wgrad_bias(const fn_conv_desc* d, const void* dY, float* dW, float* db, float* workspace, void* stream) {
  FN_REQUIRE(db != nullptr, FN_E_INVAL, "fn_conv_wgrad_bias: null db");
  return conv_wgrad_impl(d, dY, dW, db, workspace, stream);
}

static int conv_wgrad_impl(const fn_conv_desc* d, const void* dY, float* dW, float* db, float* workspace, void* stream) {
  FN_REQUIRE(d && d->x && dY && dW && workspace, FN_E_INVAL, "fn_conv_wgrad: null argument");
  FN_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, FN_E_INVAL, "fn_conv_wgrad: bad dims");
  FN_REQUIRE((d->ksize == 1 || d->ksize == 3) && (d->stride == 1 || d->stride == 2), FN_E_UNSUPPORTED,
             "fn_conv_wgrad: ksize %d stride %d", d->ksize, d->stride);
  if (d->Cin == 1 && d->prologue == FN_PRO_NONE && d->ksize == 3 && d->stride == 1 && (d->Cout & 7) == 0 && d->impl != FN_IMPL_SIMT) {
    dim3 grid(WG1_BLOCKS, d->Cout / 8);
    k_wgrad_cin1<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)d->x, (const __nv_bfloat16*)dY, workspace, d->B, d->H,
                                                          d->W, d->Cout);
    count_launch();
    int rc = check_launch("k_wgrad_cin1");
    if (rc) return rc;
    rc = reduce_partials(workspace, dW, 9ll * d->Cout, WG1_BLOCKS, (cudaStream_t)stream);
    if (rc || !db) return rc;
    return bias_grad(dY, db, workspace, (long long)d->B * d->H * d->W, d->Cout, (cudaStream_t)stream);
  }
  if (d->impl != FN_IMPL_SIMT) {      // tensor-core kernel when the shape has one
    const int OH = (d->H + d->stride - 1) / d->stride, OW = (d->W + d->stride - 1) / d->stride;
    int S = 0, rc = 0;
    float* bias_partial = nullptr;
    if (wgrad_tc_try(d->x, dY, workspace, d->B, d->H, d->W, d->Cin, OH, OW, d->Cout, d->ksize, d->stride, d->prologue, d->keep_prob,
                     d->dropout_seed, &S, (cudaStream_t)stream, rc, db != nullptr, &bias_partial,
                     (const unsigned long long*)d->dropout_seed_dev)) {
      if (rc) return rc;
      const long long nw = (long long)d->ksize * d->ksize * d->Cin * pro_mult(d->prologue) * d->Cout;
      if (db && bias_partial)        // both second stages in one launch
        return reduce_partials(workspace, dW, nw, S, (cudaStream_t)stream, bias_partial, db, d->Cout, S);
      rc = reduce_partials(workspace, dW, nw, S, (cudaStream_t)stream);
      if (rc || !db) return rc;
      return bias_grad(dY, db, workspace, (long long)d->B * OH * OW, d->Cout, (cudaStream_t)stream);
    }
    FN_REQUIRE(d->impl != FN_IMPL_TCGEN05, FN_E_UNSUPPORTED, "fn_conv_wgrad: no tcgen05 kernel for this shape");
  }
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.x = (const __nv_bfloat16*)d->x;
  p.dY = (const __nv_bfloat16*)dY;
  p.B = d->B; p.H = d->H; p.W = d->W; p.C = d->Cin;
  p.OH = (d->H + d->stride - 1) / d->stride;
  p.OW = (d->W + d->stride - 1) / d->stride;
  p.N = d->Cout; p.ksize = d->ksize; p.stride = d->stride;
  p.pad_t = same_pad_before(d->H, d->ksize, d->stride);
  p.pad_l = same_pad_before(d->W, d->ksize, d->stride);
  p.pro = d->prologue;
  p.Keff = d->Cin * pro_mult(d->prologue);
  p.transposed = 0;
  p.keep_prob = d->keep_prob;
  p.seed = d->dropout_seed;
  p.seed_dev = (const unsigned long long*)d->dropout_seed_dev;
  {
    int rc = conv_wgrad_launch(p, dW, workspace, (cudaStream_t)stream);
    if (rc || !db) return rc;
    return bias_grad(dY, db, workspace, (long long)p.B * p.OH * p.OW, p.N, (cudaStream_t)stream);
  }
}

int64_t fn_tconv_wgrad_workspace(const fn_tconv_desc* d) {
  if (!d) return 0;
  return max_ll(wgrad_workspace((long long)d->B * d->H * d->W, 9, d->Cin, d->Cout),
                wgrad_tc_workspace(d->B, 2 * d->H, 2 * d->W, d->Cout, d->H, d->W, d->Cin, 3, 2, FN_PRO_NONE));
}

/* d describes the FORWARD transposed conv (x, dims); dY: bf16 [B,2H,2W,Cout]; dW: fp32 [9, Cout, Cin] = the variable's own
 * layout [k,k,Cout,Cin] (flow_architecture.py:74).  The adjoint pairs in[i,j] with dY[2i+di, 2j+dj]: exactly the weight
 * gradient of a stride-2 conv whose input is dY and whose output gradient is x, which is how the tensor-core path runs it. */
int fn_tconv_wgrad(const fn_tconv_desc* d, const void* dY, float* dW, float* workspace, void* stream) {
  FN_REQUIRE(d && d->x && dY && dW && workspace, FN_E_INVAL, "fn_tconv_wgrad: null argument");
  FN_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, FN_E_INVAL, "fn_tconv_wgrad: bad dims");
  if (d->impl != FN_IMPL_SIMT) {
    int S = 0, rc = 0;
    if (wgrad_tc_try(dY, d->x, workspace, d->B, 2 * d->H, 2 * d->W, d->Cout, d->H, d->W, d->Cin, 3, 2, FN_PRO_NONE, 1.0f, 0, &S,
                     (cudaStream_t)stream, rc)) {
      if (rc) return rc;
      return reduce_partials(workspace, dW, 9ll * d->Cin * d->Cout, S, (cudaStream_t)stream);
    }
    FN_REQUIRE(d->impl != FN_IMPL_TCGEN05, FN_E_UNSUPPORTED, "fn_tconv_wgrad: no tcgen05 kernel for this shape");
  }
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.x = (const __nv_bfloat16*)d->x;
  p.dY = (const __nv_bfloat16*)dY;
  p.B = d->B; p.H = d->H; p.W = d->W; p.C = d->Cin;
  p.OH = 2 * d->H; p.OW = 2 * d->W;
  p.N = d->Cout; p.ksize = 3; p.stride = 2;
  p.pro = FN_PRO_NONE;
  p.Keff = d->Cin;
  p.transposed = 1;
  p.keep_prob = 1.0f;
  return conv_wgrad_launch(p, dW, workspace, (cudaStream_t)stream);
}

}  // extern "C"
