// CUDA-core (SIMT) fused convolution kernels.
//
// Role: (1) the layers of flow_net that have no tensor-core GEMM shape -- the 1-channel head
// (K = 9) and the 2-channel tanh tail (N = 2); (2) an on-device cross-check for the tcgen05
// kernels (same bf16 storage points, fp32 accumulation), selectable with FN_IMPL_SIMT.
// Semantics: tf.nn.conv2d / conv2d_transpose 'SAME' + the fused neighbours listed in
// include/flownet.h (reference: model/flow_architecture.py:57-87, 129-165).
#include "common.cuh"

namespace fn {

constexpr int SIMT_THREADS = 128;
constexpr int SIMT_NT = 8;          // output-channel slots per thread
constexpr int SIMT_MAX_KEFF = 256;  // K per tap staged in shared memory

struct ConvParams {
  fn_conv_desc d;
  int OH, OW, pad_t, pad_l;
  int Keff, Kskip_eff;
  int slots;   // output channels written (gated: Cout/2)
};

// accumulate raw channels [c0, c0+CR) of one tap for this thread's pixel; wsm holds the matching
// weight rows as [r][NACC], r < CR for the first prologue output and CR + r for the second (concat)
template <int NACC>
__device__ __forceinline__ void accumulate_pixel(const __nv_bfloat16* __restrict__ px, int Craw, int c0, int CR, int pro,
                                                 const float* __restrict__ wsm, float (&acc)[NACC],
                                                 bool drop, uint64_t seed, uint64_t base_idx, float keep) {
  const int mult = pro_mult(pro);
  const bool vec = (Craw & 7) == 0;
  for (int cb = 0; cb < CR; cb += 8) {
    bf16x8 raw;
    int nval = 8;
    if (vec) {
      raw.v = *reinterpret_cast<const uint4*>(px + c0 + cb);
    } else {
      nval = min(8, CR - cb);
      for (int j = 0; j < nval; ++j) raw.h[j] = px[c0 + cb + j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j >= nval) break;
      float p0, p1;
      pro_apply(pro, __bfloat162float(raw.h[j]), p0, p1);
      if (pro != FN_PRO_NONE) {   // the stored operand is bf16 on every path
        p0 = __bfloat162float(__float2bfloat16(p0));
        p1 = __bfloat162float(__float2bfloat16(p1));
      }
      const int cl = cb + j;      // chunk-local raw channel
      if (drop) {
        p0 *= dropout_scale(seed, base_idx + c0 + cl, keep);
        if (mult == 2) p1 *= dropout_scale(seed, base_idx + Craw + c0 + cl, keep);
      }
      const float* w0 = wsm + cl * NACC;
#pragma unroll
      for (int n = 0; n < NACC; ++n) acc[n] = fmaf(p0, w0[n], acc[n]);
      if (mult == 2) {
        const float* w1 = wsm + (CR + cl) * NACC;
#pragma unroll
        for (int n = 0; n < NACC; ++n) acc[n] = fmaf(p1, w1[n], acc[n]);
      }
    }
  }
}

// stage the weight rows of raw channels [c0, c0+CR) of one tap (w: [mult*Craw][Cout]) as fp32 [r][NACC]:
// columns n0..n0+8 and, gated, F+n0..F+n0+8
template <int NACC>
__device__ __forceinline__ void stage_weights(const __nv_bfloat16* __restrict__ w, int Craw, int c0, int CR, int mult,
                                              int Cout, int n0, int F, float* wsm) {
  for (int i = threadIdx.x; i < mult * CR * NACC; i += SIMT_THREADS) {
    const int r = i / NACC, j = i % NACC;
    const int k = (r < CR) ? (c0 + r) : (Craw + c0 + (r - CR));
    const int c = n0 + (j % SIMT_NT);            // output slot
    const int n = (j >= SIMT_NT) ? F + c : c;    // second half of NACC = the gate channel of slot c
    wsm[i] = (c < F) ? __bfloat162float(w[(size_t)k * Cout + n]) : 0.f;
  }
}

template <bool GATED>
__global__ void __launch_bounds__(SIMT_THREADS) k_conv_simt(const ConvParams p) {
  constexpr int NACC = GATED ? 2 * SIMT_NT : SIMT_NT;
  __shared__ float wsm[SIMT_MAX_KEFF * NACC];
  const fn_conv_desc& d = p.d;
  const int F = GATED ? d.Cout / 2 : d.Cout;
  const int n0 = blockIdx.y * SIMT_NT;
  const long long npix = (long long)d.B * p.OH * p.OW;
  const long long pix = (long long)blockIdx.x * SIMT_THREADS + threadIdx.x;
  const bool live = pix < npix;
  int b = 0, oy = 0, ox = 0;
  if (live) {
    b = (int)(pix / ((long long)p.OH * p.OW));
    int r = (int)(pix % ((long long)p.OH * p.OW));
    oy = r / p.OW;
    ox = r % p.OW;
  }
  float acc[NACC];
#pragma unroll
  for (int n = 0; n < NACC; ++n) acc[n] = 0.f;

  const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(d.x);
  const __nv_bfloat16* w = reinterpret_cast<const __nv_bfloat16*>(d.w);
  const bool drop = d.keep_prob < 1.0f;
  const int k = d.ksize;
  const int mult = pro_mult(d.prologue);
  const int CRmax = SIMT_MAX_KEFF / mult;       // raw channels per staged chunk
  for (int tap = 0; tap < k * k; ++tap) {
    const int di = tap / k, dj = tap % k;
    const int iy = oy * d.stride + di - p.pad_t;
    const int ix = ox * d.stride + dj - p.pad_l;
    const bool inb = live && iy >= 0 && iy < d.H && ix >= 0 && ix < d.W;
    const size_t pidx = inb ? ((size_t)b * d.H + iy) * d.W + ix : 0;
    for (int c0 = 0; c0 < d.Cin; c0 += CRmax) {
      const int CR = min(CRmax, d.Cin - c0);
      __syncthreads();
      stage_weights<NACC>(w + (size_t)tap * p.Keff * d.Cout, d.Cin, c0, CR, mult, d.Cout, n0, F, wsm);
      __syncthreads();
      if (inb)
        accumulate_pixel<NACC>(x + pidx * d.Cin, d.Cin, c0, CR, d.prologue, wsm, acc, drop,
                               d.dropout_seed + (d.dropout_seed_dev ? *d.dropout_seed_dev : 0ull),
                               d.dropout_offset + pidx * p.Keff, d.keep_prob);
    }
  }
  if (d.skip != nullptr) {
    const bool inb = live && oy < d.SH && ox < d.SW;
    const size_t pidx = inb ? ((size_t)b * d.SH + oy) * d.SW + ox : 0;
    for (int c0 = 0; c0 < d.Cskip; c0 += CRmax) {
      const int CR = min(CRmax, d.Cskip - c0);
      __syncthreads();
      stage_weights<NACC>(reinterpret_cast<const __nv_bfloat16*>(d.w_skip), d.Cskip, c0, CR, mult, d.Cout, n0, F, wsm);
      __syncthreads();
      if (inb)
        accumulate_pixel<NACC>(reinterpret_cast<const __nv_bfloat16*>(d.skip) + pidx * d.Cskip, d.Cskip, c0, CR,
                               d.prologue, wsm, acc, false, 0, 0, 1.f);
    }
  }
  if (!live) return;

  // ---- epilogue
  float out[SIMT_NT];
#pragma unroll
  for (int j = 0; j < SIMT_NT; ++j) {
    const int n = n0 + j;
    float v = 0.f;
    if (n < F) {
      if (GATED) {
        v = (acc[j] + d.bias[n]) * sigmoid_f(acc[SIMT_NT + j] + d.bias[F + n]);
      } else {
        v = acc[j] + d.bias[n];
      }
      if (d.epilogue == FN_EPI_GATE_RES || d.epilogue == FN_EPI_RES) {
        const int cr = n - (F - d.Cres);          // channels are zero-padded IN FRONT
        if (d.res != nullptr && cr >= 0) {
          const __nv_bfloat16* res = reinterpret_cast<const __nv_bfloat16*>(d.res);
          if (d.res_pool) {
            float s = 0.f;
            int cnt = 0;
            for (int a = 0; a < 2; ++a)
              for (int c = 0; c < 2; ++c) {
                const int ry = 2 * oy + a, rx = 2 * ox + c;
                if (ry < d.RH && rx < d.RW) {
                  s += __bfloat162float(res[(((size_t)b * d.RH + ry) * d.RW + rx) * d.Cres + cr]);
                  ++cnt;
                }
              }
            v += s / (float)cnt;
          } else {
            v += __bfloat162float(res[(((size_t)b * d.RH + oy) * d.RW + ox) * d.Cres + cr]);
          }
        }
      }
      if (d.epilogue == FN_EPI_TANH) v = tanhf(v);
    }
    out[j] = v;
  }
  const size_t obase = (size_t)pix * F + n0;
  if (d.epilogue == FN_EPI_TANH) {
    float* y = reinterpret_cast<float*>(d.y);
    for (int j = 0; j < SIMT_NT && n0 + j < F; ++j) y[obase + j] = out[j];
  } else {
    __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(d.y);
    if ((F & 7) == 0) {
      bf16x8 o;
#pragma unroll
      for (int j = 0; j < SIMT_NT; ++j) o.h[j] = __float2bfloat16(out[j]);
      *reinterpret_cast<uint4*>(y + obase) = o.v;
    } else {
      for (int j = 0; j < SIMT_NT && n0 + j < F; ++j) y[obase + j] = __float2bfloat16(out[j]);
    }
  }
}

// ------------------------------------------------------------------ head / tail of the network
// The 1-channel head (boundary -> F, K = 9) and the 2-channel tanh tail (F -> 2) are pure HBM
// streams with no GEMM shape: one thread per pixel, every output channel in registers, weights as
// fp32 broadcast reads from shared memory, 3x3 neighbourhood through L1.
// reference: first res_block conv_1 (flow_architecture.py:132-133) and last_conv + tanh (:242-243).
template <int CIN, int NOUT, bool TANH>
__global__ void __launch_bounds__(256) k_direct3x3(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                                   const float* __restrict__ bias, void* __restrict__ y, int B, int H,
                                                   int W) {
  __shared__ float ws[9 * CIN * NOUT + NOUT];
  for (int i = threadIdx.x; i < 9 * CIN * NOUT; i += 256) ws[i] = __bfloat162float(w[i]);
  for (int i = threadIdx.x; i < NOUT; i += 256) ws[9 * CIN * NOUT + i] = bias[i];
  __syncthreads();
  const long long npix = (long long)B * H * W;
  for (long long pix = (long long)blockIdx.x * 256 + threadIdx.x; pix < npix; pix += (long long)gridDim.x * 256) {
    const int ox = (int)(pix % W);
    const int oy = (int)((pix / W) % H);
    float acc[NOUT];
#pragma unroll
    for (int n = 0; n < NOUT; ++n) acc[n] = ws[9 * CIN * NOUT + n];
#pragma unroll
    for (int di = 0; di < 3; ++di) {
      const int iy = oy + di - 1;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int dj = 0; dj < 3; ++dj) {
        const int ix = ox + dj - 1;
        if (ix < 0 || ix >= W) continue;
        const __nv_bfloat16* px = x + (pix + (long long)(di - 1) * W + (dj - 1)) * CIN;
        const float* wt = ws + (di * 3 + dj) * CIN * NOUT;
        if (CIN == 1) {
          const float v = __bfloat162float(px[0]);
#pragma unroll
          for (int n = 0; n < NOUT; ++n) acc[n] = fmaf(v, wt[n], acc[n]);
        } else {
#pragma unroll
          for (int c8 = 0; c8 < CIN; c8 += 8) {
            bf16x8 raw;
            raw.v = __ldg(reinterpret_cast<const uint4*>(px + c8));
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float v = __bfloat162float(raw.h[j]);
#pragma unroll
              for (int n = 0; n < NOUT; ++n) acc[n] = fmaf(v, wt[(c8 + j) * NOUT + n], acc[n]);
            }
          }
        }
      }
    }
    if (TANH) {
      float* yo = reinterpret_cast<float*>(y) + pix * NOUT;
      if (NOUT == 2) {
        *reinterpret_cast<float2*>(yo) = make_float2(tanhf(acc[0]), tanhf(acc[1]));
      } else {
#pragma unroll
        for (int n = 0; n < NOUT; ++n) yo[n] = tanhf(acc[n]);
      }
    } else {
      __nv_bfloat16* yo = reinterpret_cast<__nv_bfloat16*>(y) + pix * NOUT;
#pragma unroll
      for (int n0 = 0; n0 < NOUT; n0 += 8) {
        bf16x8 o;
#pragma unroll
        for (int j = 0; j < 8; ++j) o.h[j] = __float2bfloat16(acc[n0 + j]);
        *reinterpret_cast<uint4*>(yo + n0) = o.v;
      }
    }
  }
}

template <int CIN, int NOUT, bool TANH>
static int launch_direct(const fn_conv_desc& d, cudaStream_t stream) {
  const long long npix = (long long)d.B * d.H * d.W;
  long long grid = (npix + 255) / 256;
  if (grid > 148LL * 16) grid = 148LL * 16;
  k_direct3x3<CIN, NOUT, TANH><<<(unsigned)grid, 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(d.x), reinterpret_cast<const __nv_bfloat16*>(d.w), d.bias, d.y, d.B, d.H, d.W);
  count_launch();
  return check_launch("k_direct3x3");
}

// Head of the network, 1 channel -> 8 (first res_block conv_1, flow_architecture.py:132-133 on the boundary): a CTA
// stages an 8 x 128 pixel tile (+halo) of the boundary image as fp32 in shared memory, all loads issued before the
// first use.  A thread produces the pixels lane, lane+32, lane+64, lane+96 of one tile row, so shared-memory reads are
// conflict-free and every warp store covers 512 contiguous bytes.
template <int NOUT>
__global__ void __launch_bounds__(256) k_head3x3(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                                    const float* __restrict__ bias, __nv_bfloat16* __restrict__ y,
                                                    __nv_bfloat16* __restrict__ yp, int B, int H, int W, int tiles_x,
                                                    int tiles_y) {
  __shared__ __align__(16) float ws[9 * NOUT + NOUT];
  __shared__ float tile[10][132];
  if (threadIdx.x < 9 * NOUT) ws[threadIdx.x] = __bfloat162float(w[threadIdx.x]);
  if (threadIdx.x < NOUT) ws[9 * NOUT + threadIdx.x] = bias[threadIdx.x];
  int t = blockIdx.x;
  const int tx = t % tiles_x;
  t /= tiles_x;
  const int ty = t % tiles_y;
  const int b = t / tiles_y;
  const int y0 = ty * 8, x0 = tx * 128;
  {
    float v[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const int i = threadIdx.x + 256 * k;
      const int r = i / 130, c = i - r * 130;
      const int iy = y0 + r - 1, ix = x0 + c - 1;
      v[k] = 0.f;
      if (i < 1300 && iy >= 0 && iy < H && ix >= 0 && ix < W) v[k] = __bfloat162float(x[((size_t)b * H + iy) * W + ix]);
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const int i = threadIdx.x + 256 * k;
      const int r = i / 130, c = i - r * 130;
      if (i < 1300) tile[r][c] = v[k];
    }
  }
  __syncthreads();
  const int row = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int oy = y0 + row;
  if (oy >= H) return;
  float acc[4][NOUT];
#pragma unroll
  for (int px = 0; px < 4; ++px)
#pragma unroll
    for (int n = 0; n < NOUT; ++n) acc[px][n] = ws[9 * NOUT + n];
#pragma unroll
  for (int di = 0; di < 3; ++di)
#pragma unroll
    for (int dj = 0; dj < 3; ++dj) {
#pragma unroll
      for (int n4 = 0; n4 < NOUT; n4 += 4) {
        const float4 wv = *reinterpret_cast<const float4*>(&ws[(di * 3 + dj) * NOUT + n4]);
#pragma unroll
        for (int px = 0; px < 4; ++px) {
          const float v = tile[row + di][lane + 32 * px + dj];
          acc[px][n4] = fmaf(v, wv.x, acc[px][n4]); acc[px][n4 + 1] = fmaf(v, wv.y, acc[px][n4 + 1]);
          acc[px][n4 + 2] = fmaf(v, wv.z, acc[px][n4 + 2]); acc[px][n4 + 3] = fmaf(v, wv.w, acc[px][n4 + 3]);
        }
      }
    }
  __nv_bfloat16* yo = y + (((size_t)b * H + oy) * W + x0) * NOUT;
#pragma unroll
  for (int px = 0; px < 4; ++px) {
    const int ox = lane + 32 * px;
    if (x0 + ox < W) {
#pragma unroll
      for (int n0 = 0; n0 < NOUT; n0 += 8) {
        bf16x8 o;
#pragma unroll
        for (int n = 0; n < 8; ++n) o.h[n] = __float2bfloat16(acc[px][n0 + n]);
        if (y) *reinterpret_cast<uint4*>(yo + (size_t)ox * NOUT + n0) = o.v;
        if (yp) {    // planes form of the output (fn_conv_desc.y_planes): [B][2 NOUT/8][H][W][8] = concat_elu(y)
          uint4 e0, e1;
          concat_elu8(o.v, e0, e1);
          __nv_bfloat16* d0 = yp + ((((size_t)b * (NOUT / 4) + n0 / 8) * H + oy) * W + x0 + ox) * 8;
          *reinterpret_cast<uint4*>(d0) = e0;
          *reinterpret_cast<uint4*>(d0 + (size_t)(NOUT / 8) * H * W * 8) = e1;
        }
      }
    }
  }
}

// Round-2 head for 8 output channels: 16 x 64 tile, a thread owns a vertical strip of 4 output pixels and loads its 6 x 3
// input values once (18 shared loads per 4 pixels instead of 36), 288 FMAs per strip; stores as in k_head3x3.
__device__ __forceinline__ float head_in(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float head_in(unsigned char v) { return (float)v; }
// TIN = unsigned char: the boundary image as it arrives (fn_conv_desc.x_u8) -- no separate cast launch
template <typename TIN>
__global__ void __launch_bounds__(256) k_head8_v2(const TIN* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                                  const float* __restrict__ bias, __nv_bfloat16* __restrict__ y,
                                                  __nv_bfloat16* __restrict__ yp, int B, int H, int W, int tiles_x, int tiles_y) {
  constexpr int TW = 64, TH = 16, PW = TW + 2, PH = TH + 2;
  __shared__ __align__(16) float ws[9 * 8 + 8];
  __shared__ float tile[PH][PW + 2];
  if (threadIdx.x < 72) ws[threadIdx.x] = __bfloat162float(w[threadIdx.x]);
  if (threadIdx.x < 8) ws[72 + threadIdx.x] = bias[threadIdx.x];
  int t = blockIdx.x;
  const int tx = t % tiles_x;
  t /= tiles_x;
  const int ty = t % tiles_y;
  const int b = t / tiles_y;
  const int y0 = ty * TH, x0 = tx * TW;
  {
    constexpr int NPX = PH * PW, NLD = (NPX + 255) / 256;
    float v[NLD];
#pragma unroll
    for (int k = 0; k < NLD; ++k) {
      const int i = threadIdx.x + 256 * k;
      const int r = i / PW, c = i - r * PW;
      const int iy = y0 + r - 1, ix = x0 + c - 1;
      v[k] = 0.f;
      if (i < NPX && iy >= 0 && iy < H && ix >= 0 && ix < W) v[k] = head_in(x[((size_t)b * H + iy) * W + ix]);
    }
#pragma unroll
    for (int k = 0; k < NLD; ++k) {
      const int i = threadIdx.x + 256 * k;
      const int r = i / PW, c = i - r * PW;
      if (i < NPX) tile[r][c] = v[k];
    }
  }
  __syncthreads();
  const int cx = threadIdx.x & 63, strip = threadIdx.x >> 6;
  float acc[4][8];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int n = 0; n < 8; ++n) acc[r][n] = ws[72 + n];
#pragma unroll
  for (int dj = 0; dj < 3; ++dj) {
    float in[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) in[r] = tile[strip * 4 + r][cx + dj];
#pragma unroll
    for (int di = 0; di < 3; ++di) {
      const float4 w0 = *reinterpret_cast<const float4*>(&ws[(di * 3 + dj) * 8]);
      const float4 w1 = *reinterpret_cast<const float4*>(&ws[(di * 3 + dj) * 8 + 4]);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float v = in[r + di];
        acc[r][0] = fmaf(v, w0.x, acc[r][0]); acc[r][1] = fmaf(v, w0.y, acc[r][1]);
        acc[r][2] = fmaf(v, w0.z, acc[r][2]); acc[r][3] = fmaf(v, w0.w, acc[r][3]);
        acc[r][4] = fmaf(v, w1.x, acc[r][4]); acc[r][5] = fmaf(v, w1.y, acc[r][5]);
        acc[r][6] = fmaf(v, w1.z, acc[r][6]); acc[r][7] = fmaf(v, w1.w, acc[r][7]);
      }
    }
  }
  const int ox = x0 + cx;
  if (ox >= W) return;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int oy = y0 + strip * 4 + r;
    if (oy >= H) break;
    bf16x8 o;
#pragma unroll
    for (int n = 0; n < 8; ++n) o.h[n] = __float2bfloat16(acc[r][n]);
    const size_t pix = ((size_t)b * H + oy) * W + ox;
    if (y) *reinterpret_cast<uint4*>(y + pix * 8) = o.v;
    if (yp) {    // planes form of the output (fn_conv_desc.y_planes): [B][2][H][W][8] = concat_elu(y)
      uint4 e0, e1;
      concat_elu8(o.v, e0, e1);
      __nv_bfloat16* d0 = yp + ((((size_t)b * 2) * H + oy) * W + ox) * 8;
      *reinterpret_cast<uint4*>(d0) = e0;
      *reinterpret_cast<uint4*>(d0 + (size_t)H * W * 8) = e1;
    }
  }
}

static int g_head_v2 = 1;      // debug / A-B: 0 sends the 8-channel head back to k_head3x3

static int launch_head8_v2(const fn_conv_desc& d, cudaStream_t stream) {
  const int tiles_x = (d.W + 63) / 64, tiles_y = (d.H + 15) / 16;
  const long long grid = (long long)d.B * tiles_x * tiles_y;
  if (d.x_u8)
    k_head8_v2<unsigned char><<<(unsigned)grid, 256, 0, stream>>>(reinterpret_cast<const unsigned char*>(d.x), reinterpret_cast<const __nv_bfloat16*>(d.w),
                                                                  d.bias, reinterpret_cast<__nv_bfloat16*>(d.y), reinterpret_cast<__nv_bfloat16*>(d.y_planes),
                                                                  d.B, d.H, d.W, tiles_x, tiles_y);
  else
    k_head8_v2<__nv_bfloat16><<<(unsigned)grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(d.x), reinterpret_cast<const __nv_bfloat16*>(d.w),
                                                                  d.bias, reinterpret_cast<__nv_bfloat16*>(d.y), reinterpret_cast<__nv_bfloat16*>(d.y_planes),
                                                                  d.B, d.H, d.W, tiles_x, tiles_y);
  count_launch();
  return check_launch("k_head8_v2");
}

template <int NOUT>
static int launch_head(const fn_conv_desc& d, cudaStream_t stream) {
  const int tiles_x = (d.W + 127) / 128, tiles_y = (d.H + 7) / 8;
  const long long grid = (long long)d.B * tiles_x * tiles_y;
  k_head3x3<NOUT><<<(unsigned)grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(d.x),
                                                   reinterpret_cast<const __nv_bfloat16*>(d.w), d.bias,
                                                   reinterpret_cast<__nv_bfloat16*>(d.y),
                                                   reinterpret_cast<__nv_bfloat16*>(d.y_planes), d.B, d.H, d.W, tiles_x, tiles_y);
  count_launch();
  return check_launch("k_head3x3");
}

// Tail of the network, 8 channels -> 2 + tanh, fp32 out (last_conv, flow_architecture.py:242-243): the same tiling;
// the 8 x 128 (+halo) tile is staged as it is in memory (one 16-byte bf16 vector per pixel), a thread owns four pixels
// 32 apart, reads each tap's 16 weights once for the four of them, and writes float2 per pixel (256 contiguous bytes
// per warp store).  A third of the instructions of the one-thread-per-pixel kernel above.
template <int CIN>
__global__ void __launch_bounds__(256) k_tail3x3(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                                    const float* __restrict__ bias, float* __restrict__ y, int B, int H, int W,
                                                    int tiles_x, int tiles_y) {
  constexpr int CH = CIN / 8;
  __shared__ __align__(16) float ws[18 * CIN + 4];
  __shared__ uint4 tile[10][130][CH];
  for (int i = threadIdx.x; i < 18 * CIN; i += 256) ws[i] = __bfloat162float(w[i]);        // [tap][c][n]
  if (threadIdx.x < 2) ws[18 * CIN + threadIdx.x] = bias[threadIdx.x];
  int t = blockIdx.x;
  const int tx = t % tiles_x;
  t /= tiles_x;
  const int ty = t % tiles_y;
  const int b = t / tiles_y;
  const int y0 = ty * 8, x0 = tx * 128;
  for (int q = 0; q < CH; ++q) {
    uint4 v[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const int i = threadIdx.x + 256 * k;
      const int r = i / 130, c = i - r * 130;
      const int iy = y0 + r - 1, ix = x0 + c - 1;
      v[k] = make_uint4(0u, 0u, 0u, 0u);
      if (i < 1300 && iy >= 0 && iy < H && ix >= 0 && ix < W)
        v[k] = __ldg(reinterpret_cast<const uint4*>(x + (((size_t)b * H + iy) * W + ix) * CIN + q * 8));
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const int i = threadIdx.x + 256 * k;
      const int r = i / 130, c = i - r * 130;
      if (i < 1300) tile[r][c][q] = v[k];
    }
  }
  __syncthreads();
  const int row = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int oy = y0 + row;
  if (oy >= H) return;
  float a0[4], a1[4];
#pragma unroll
  for (int px = 0; px < 4; ++px) { a0[px] = ws[18 * CIN]; a1[px] = ws[18 * CIN + 1]; }
#pragma unroll
  for (int di = 0; di < 3; ++di)
#pragma unroll
    for (int dj = 0; dj < 3; ++dj) {
#pragma unroll
      for (int q = 0; q < CH; ++q) {
        float wt[16];
#pragma unroll
        for (int f = 0; f < 4; ++f) {
          const float4 wv = *reinterpret_cast<const float4*>(&ws[((di * 3 + dj) * CIN + q * 8) * 2 + 4 * f]);
          wt[4 * f] = wv.x; wt[4 * f + 1] = wv.y; wt[4 * f + 2] = wv.z; wt[4 * f + 3] = wv.w;
        }
#pragma unroll
        for (int px = 0; px < 4; ++px) {
          bf16x8 in;
          in.v = tile[row + di][lane + 32 * px + dj][q];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float v = __bfloat162float(in.h[c]);
            a0[px] = fmaf(v, wt[2 * c], a0[px]);
            a1[px] = fmaf(v, wt[2 * c + 1], a1[px]);
          }
        }
      }
    }
  float* yo = y + (((size_t)b * H + oy) * W + x0) * 2;
#pragma unroll
  for (int px = 0; px < 4; ++px) {
    const int ox = lane + 32 * px;
    if (x0 + ox < W) *reinterpret_cast<float2*>(yo + (size_t)ox * 2) = make_float2(tanhf(a0[px]), tanhf(a1[px]));
  }
}

// Round-2 tail for 8 channels.  k_tail3x3 above is instruction-bound (33 us at batch 64 against a 7.7 us HBM floor): per
// pixel it issues 9 shared loads, 72 bf16 -> fp32 unpacks and 144 FMAs.  Here the 16 x 64 (+halo) tile is staged as FP32 in
// two 4-channel planes (converted once per pixel; consecutive lanes read consecutive 16-byte vectors: conflict-free) and a
// thread owns a vertical strip of 4 output pixels: the 6 x 3 input pixels of the strip are loaded once and reused by the
// three filter rows, so a pixel costs 9 shared loads of data + 9 of weights per FOUR pixels, no unpacks, 144 FMAs.
__global__ void __launch_bounds__(256) k_tail8_v2(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                                  const float* __restrict__ bias, float* __restrict__ y, int B, int H, int W,
                                                  int tiles_x, int tiles_y) {
  constexpr int TW = 64, TH = 16, PW = TW + 2, PH = TH + 2;
  __shared__ __align__(16) float ws[9 * 16 + 4];               // [tap][c][n], bias
  __shared__ float4 tile[2][PH][PW];                           // [channel half][row][col]
  for (int i = threadIdx.x; i < 144; i += 256) ws[i] = __bfloat162float(w[i]);
  if (threadIdx.x < 2) ws[144 + threadIdx.x] = bias[threadIdx.x];
  int t = blockIdx.x;
  const int tx = t % tiles_x;
  t /= tiles_x;
  const int ty = t % tiles_y;
  const int b = t / tiles_y;
  const int y0 = ty * TH, x0 = tx * TW;
  {
    constexpr int NPX = PH * PW, NLD = (NPX + 255) / 256;      // 1188 pixels, 5 per thread
    uint4 v[NLD];
#pragma unroll
    for (int k = 0; k < NLD; ++k) {
      const int i = threadIdx.x + 256 * k;
      const int r = i / PW, c = i - r * PW;
      const int iy = y0 + r - 1, ix = x0 + c - 1;
      v[k] = make_uint4(0u, 0u, 0u, 0u);
      if (i < NPX && iy >= 0 && iy < H && ix >= 0 && ix < W)
        v[k] = __ldg(reinterpret_cast<const uint4*>(x + (((size_t)b * H + iy) * W + ix) * 8));
    }
#pragma unroll
    for (int k = 0; k < NLD; ++k) {
      const int i = threadIdx.x + 256 * k;
      const int r = i / PW, c = i - r * PW;
      if (i < NPX) {
        tile[0][r][c] = make_float4(__uint_as_float(v[k].x << 16), __uint_as_float(v[k].x & 0xffff0000u),
                                    __uint_as_float(v[k].y << 16), __uint_as_float(v[k].y & 0xffff0000u));
        tile[1][r][c] = make_float4(__uint_as_float(v[k].z << 16), __uint_as_float(v[k].z & 0xffff0000u),
                                    __uint_as_float(v[k].w << 16), __uint_as_float(v[k].w & 0xffff0000u));
      }
    }
  }
  __syncthreads();
  const int cx = threadIdx.x & 63, strip = threadIdx.x >> 6;   // column, 4-row strip
  float a0[4], a1[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) { a0[r] = ws[144]; a1[r] = ws[145]; }
#pragma unroll
  for (int dj = 0; dj < 3; ++dj) {
    float4 lo[6], hi[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      lo[r] = tile[0][strip * 4 + r][cx + dj];
      hi[r] = tile[1][strip * 4 + r][cx + dj];
    }
#pragma unroll
    for (int di = 0; di < 3; ++di) {
      float wt[16];
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        const float4 wv = *reinterpret_cast<const float4*>(&ws[(di * 3 + dj) * 16 + 4 * f]);
        wt[4 * f] = wv.x; wt[4 * f + 1] = wv.y; wt[4 * f + 2] = wv.z; wt[4 * f + 3] = wv.w;
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float4 p0 = lo[r + di], p1 = hi[r + di];
        a0[r] = fmaf(p0.x, wt[0], a0[r]);  a1[r] = fmaf(p0.x, wt[1], a1[r]);
        a0[r] = fmaf(p0.y, wt[2], a0[r]);  a1[r] = fmaf(p0.y, wt[3], a1[r]);
        a0[r] = fmaf(p0.z, wt[4], a0[r]);  a1[r] = fmaf(p0.z, wt[5], a1[r]);
        a0[r] = fmaf(p0.w, wt[6], a0[r]);  a1[r] = fmaf(p0.w, wt[7], a1[r]);
        a0[r] = fmaf(p1.x, wt[8], a0[r]);  a1[r] = fmaf(p1.x, wt[9], a1[r]);
        a0[r] = fmaf(p1.y, wt[10], a0[r]); a1[r] = fmaf(p1.y, wt[11], a1[r]);
        a0[r] = fmaf(p1.z, wt[12], a0[r]); a1[r] = fmaf(p1.z, wt[13], a1[r]);
        a0[r] = fmaf(p1.w, wt[14], a0[r]); a1[r] = fmaf(p1.w, wt[15], a1[r]);
      }
    }
  }
  const int ox = x0 + cx;
  if (ox < W) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int oy = y0 + strip * 4 + r;
      if (oy < H) *reinterpret_cast<float2*>(y + (((size_t)b * H + oy) * W + ox) * 2) = make_float2(tanhf(a0[r]), tanhf(a1[r]));
    }
  }
}

static int g_tail_v2 = 1;      // debug / A-B: 0 sends the 8-channel tail back to k_tail3x3

static int launch_tail8_v2(const fn_conv_desc& d, cudaStream_t stream) {
  const int tiles_x = (d.W + 63) / 64, tiles_y = (d.H + 15) / 16;
  const long long grid = (long long)d.B * tiles_x * tiles_y;
  k_tail8_v2<<<(unsigned)grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(d.x), reinterpret_cast<const __nv_bfloat16*>(d.w),
                                                 d.bias, reinterpret_cast<float*>(d.y), d.B, d.H, d.W, tiles_x, tiles_y);
  count_launch();
  return check_launch("k_tail8_v2");
}

template <int CIN>
static int launch_tail(const fn_conv_desc& d, cudaStream_t stream) {
  const int tiles_x = (d.W + 127) / 128, tiles_y = (d.H + 7) / 8;
  const long long grid = (long long)d.B * tiles_x * tiles_y;
  k_tail3x3<CIN><<<(unsigned)grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(d.x),
                                                   reinterpret_cast<const __nv_bfloat16*>(d.w), d.bias,
                                                   reinterpret_cast<float*>(d.y), d.B, d.H, d.W, tiles_x, tiles_y);
  count_launch();
  return check_launch("k_tail3x3");
}

// returns true (and sets rc) if a specialised head/tail kernel took the launch
static bool try_direct(const fn_conv_desc& d, cudaStream_t stream, int& rc) {
  if (d.ksize != 3 || d.stride != 1 || d.prologue != FN_PRO_NONE || d.skip || d.keep_prob < 1.0f) return false;
  if (d.epilogue == FN_EPI_BIAS && d.Cin == 1 && d.Cout == 8) { rc = (g_head_v2 || d.x_u8) ? launch_head8_v2(d, stream) : launch_head<8>(d, stream); return true; }
  if (d.epilogue == FN_EPI_BIAS && d.Cin == 1 && d.Cout == 16) { rc = launch_head<16>(d, stream); return true; }
  if (d.epilogue == FN_EPI_BIAS && d.Cin == 1) {
    if (d.Cout == 8) { rc = launch_direct<1, 8, false>(d, stream); return true; }
    if (d.Cout == 16) { rc = launch_direct<1, 16, false>(d, stream); return true; }
    if (d.Cout == 32) { rc = launch_direct<1, 32, false>(d, stream); return true; }
  }
  if (d.epilogue == FN_EPI_TANH && d.Cout == 2) {
    if (d.Cin == 8) { rc = g_tail_v2 ? launch_tail8_v2(d, stream) : launch_tail<8>(d, stream); return true; }
    if (d.Cin == 16) { rc = launch_tail<16>(d, stream); return true; }
    if (d.Cin == 32) { rc = launch_direct<32, 2, true>(d, stream); return true; }
  }
  return false;
}

int conv_fwd_simt(const fn_conv_desc& d, cudaStream_t stream) {
  int rc_direct = 0;
  if (try_direct(d, stream, rc_direct)) return rc_direct;
  ConvParams p;
  p.d = d;
  p.OH = (d.H + d.stride - 1) / d.stride;
  p.OW = (d.W + d.stride - 1) / d.stride;
  p.pad_t = same_pad_before(d.H, d.ksize, d.stride);
  p.pad_l = same_pad_before(d.W, d.ksize, d.stride);
  p.Keff = d.Cin * pro_mult(d.prologue);
  p.Kskip_eff = d.Cskip * pro_mult(d.prologue);
  const bool gated = d.epilogue == FN_EPI_GATE_RES;
  p.slots = gated ? d.Cout / 2 : d.Cout;
  const long long npix = (long long)d.B * p.OH * p.OW;
  dim3 grid((unsigned)((npix + SIMT_THREADS - 1) / SIMT_THREADS), (unsigned)((p.slots + SIMT_NT - 1) / SIMT_NT));
  if (gated)
    k_conv_simt<true><<<grid, SIMT_THREADS, 0, stream>>>(p);
  else
    k_conv_simt<false><<<grid, SIMT_THREADS, 0, stream>>>(p);
  count_launch();
  return check_launch("k_conv_simt");
}

// ------------------------------------------------------------------ transposed conv (k=3, s=2)
// out[b, I, J, o] = bias[o] + sum_{di,dj : I-di, J-dj even, in range} x[b,(I-di)/2,(J-dj)/2,:] . W[di,dj,:,o]
__global__ void __launch_bounds__(SIMT_THREADS) k_tconv_simt(const fn_tconv_desc d) {
  __shared__ float wsm[SIMT_MAX_KEFF * SIMT_NT];
  const int OH = 2 * d.H, OW = 2 * d.W;
  const int n0 = blockIdx.y * SIMT_NT;
  const long long npix = (long long)d.B * OH * OW;
  const long long pix = (long long)blockIdx.x * SIMT_THREADS + threadIdx.x;
  const bool live = pix < npix;
  int b = 0, I = 0, J = 0;
  if (live) {
    b = (int)(pix / ((long long)OH * OW));
    int r = (int)(pix % ((long long)OH * OW));
    I = r / OW;
    J = r % OW;
  }
  float acc[SIMT_NT];
#pragma unroll
  for (int n = 0; n < SIMT_NT; ++n) acc[n] = 0.f;
  const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(d.x);
  const __nv_bfloat16* w = reinterpret_cast<const __nv_bfloat16*>(d.w);
  for (int tap = 0; tap < 9; ++tap) {
    const int di = tap / 3, dj = tap % 3;
    const int ti = I - di, tj = J - dj;
    const bool inb = live && ti >= 0 && tj >= 0 && ((ti | tj) & 1) == 0 && (ti >> 1) < d.H && (tj >> 1) < d.W;
    const size_t pidx = inb ? ((size_t)b * d.H + (ti >> 1)) * d.W + (tj >> 1) : 0;
    for (int c0 = 0; c0 < d.Cin; c0 += SIMT_MAX_KEFF) {
      const int CR = min(SIMT_MAX_KEFF, d.Cin - c0);
      __syncthreads();
      stage_weights<SIMT_NT>(w + (size_t)tap * d.Cin * d.Cout, d.Cin, c0, CR, 1, d.Cout, n0, d.Cout, wsm);
      __syncthreads();
      if (inb) accumulate_pixel<SIMT_NT>(x + pidx * d.Cin, d.Cin, c0, CR, FN_PRO_NONE, wsm, acc, false, 0, 0, 1.f);
    }
  }
  if (!live) return;
  __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(d.y);
  bf16x8 o;
#pragma unroll
  for (int j = 0; j < SIMT_NT; ++j) o.h[j] = __float2bfloat16(acc[j] + d.bias[n0 + j]);
  *reinterpret_cast<uint4*>(y + (size_t)pix * d.Cout + n0) = o.v;
}

int tconv_fwd_simt(const fn_tconv_desc& d, cudaStream_t stream) {
  const long long npix = (long long)d.B * 4 * d.H * d.W;
  dim3 grid((unsigned)((npix + SIMT_THREADS - 1) / SIMT_THREADS), (unsigned)(d.Cout / SIMT_NT));
  k_tconv_simt<<<grid, SIMT_THREADS, 0, stream>>>(d);
  count_launch();
  return check_launch("k_tconv_simt");
}

}  // namespace fn

extern "C" {
void fn_debug_set_tail_v2(int v) { fn::g_tail_v2 = v ? 1 : 0; fn::g_head_v2 = v ? 1 : 0; }
}
