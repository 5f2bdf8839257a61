// Template-specialised tcgen05 kernel for the DEEP stride-1 3x3 convolutions of flow_net
// (concat_elu prologue, C in {64,128}: the 16x32 and 8x16 levels).  These launches are tensor-bound
// in theory and weight-streaming-bound in practice (295 KB - 1.18 MB of weights per launch against
// 128-512 pixels per image), with only one or two tiles per SM, so the kernel is organised around
// per-phase latency rather than cross-tile pipelining:
//
//   warp 8      producer   starts streaming the packed weights through a 2-3 x 32 KB ring (cp.async.bulk
//                          + mbarrier) at kernel entry, TMA-loads the raw halo tile / skip tile (tensor
//                          map, hardware zero fill) and, once the transform has consumed the stage
//                          buffer, the shortcut tile into the same shared memory
//   warps 0-7   workers    all 256 threads transform stage -> concat_elu A planes; warps 0..J-1 elect
//                          one lane each to issue the fully unrolled tcgen05.mma program (chunk waits
//                          at compile-time positions); all 8 warps run the epilogue, each on one TMEM
//                          lane quadrant and one half of the channels
//
// The 8x16 level uses the transposed tile orientation (8 pixels along y per UMMA row group) so that a
// whole image is one M=128 tile.  Same math, layouts and MMA order as k_conv_tc / k_conv_s1p.
// reference semantics: res_block conv_1 (+nin) and conv_2 (+gate, +shortcut), flow_architecture.py:129-165
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"
#include "tma.cuh"

namespace fn {
using namespace tma;
namespace s1d {

constexpr int THREADS = 288;
constexpr int HDR = 256 + 1024;
constexpr int CHUNK_BYTES = 32768;

struct Params {
  CUtensorMap tm_x, tm_skip, tm_res;
  CUtensorMap tm_w;                    // NSPLIT = 2: the packed weights as 8-byte units (2 FL, parts, 2 K halves, k-steps)
  const __nv_bfloat16* wpk;
  const float* bias;
  __nv_bfloat16* y;
  const __nv_bfloat16* res;
  int* status;
  int B, H, W;
  int tiles_u, tiles_v, ntiles;
  int Cout, res_mode, Cres, RH, RW, res_pool;
  int gate_bwd;                        // FN_EPI_GATE_BWD (see conv_s1p.cu)
  float keep_prob;                     // < 1: dropout on the prologue output (training)
  unsigned long long dropout_seed;
  const unsigned long long* dropout_seed_dev;
};

// NSPLIT = 2: two CTAs share a tile, each computes half of the output channels (for the gate: the u and the v columns of
// the same channels) and streams only its half of the weights -- the 8x16 level at batch 64 is 64 tiles for 148 SMs
template <int CLOG, int J, bool GATE, bool SKIP, bool TR, int NST, int NSPLIT = 1>
struct Cfg {
  static constexpr int CH = 1 << CLOG, C = 8 * CH, F = C;
  static constexpr int N = GATE ? 2 * C : C;
  static constexpr int FL = F / NSPLIT;            // output channels of one CTA
  static constexpr int NL = N / NSPLIT;            // accumulator columns of one CTA (u part, then v part)
  static constexpr int PARTS = GATE ? 2 : 1;
  static constexpr int KS = CH;
  static constexpr int KSTEPS = 9 * KS + (SKIP ? KS : 0);
  static constexpr int PU = 8 * J + 2, PV = 18, P = PU * PV;
  static constexpr int PPAD = P + ((1 - (P & 7)) + 8) % 8;          // CH >= 8: stride = 1 (mod 8) pixels
  static constexpr int SPC = CHUNK_BYTES / (NL * 32);                // k-steps per weight chunk
  static constexpr int NCHUNKS = (KSTEPS + SPC - 1) / SPC;
  static constexpr int PLANE_BYTES = (2 * CH * PPAD * 16 + 127) / 128 * 128;
  static constexpr int STAGE_BYTES = ((P * C * 2) + 127) / 128 * 128;
  static constexpr int RES_BYTES = GATE ? 128 * J * F * 2 : 0;
  static constexpr int STAGE_REGION = STAGE_BYTES > RES_BYTES ? STAGE_BYTES : RES_BYTES;   // shortcut aliases the stage
  static constexpr int OFF_PLANES = HDR;
  static constexpr int OFF_PLANES_S = OFF_PLANES + PLANE_BYTES;
  static constexpr int OFF_STAGE = OFF_PLANES_S + (SKIP ? PLANE_BYTES : 0);
  static constexpr int OFF_STAGE_S = OFF_STAGE + STAGE_REGION;
  static constexpr int OFF_RING = OFF_STAGE_S + (SKIP ? STAGE_BYTES : 0);
  static constexpr int SMEM = OFF_RING + NST * CHUNK_BYTES + 128;
  static constexpr int ACC_COLS = J * NL;
  static constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128 : ACC_COLS <= 256 ? 256 : 512;
  static_assert(ACC_COLS <= 512, "accumulators exceed TMEM");
  static_assert(SPC >= 1, "chunk smaller than one k-step");
  static_assert(NSPLIT == 1 || (!SKIP && FL % 16 == 0 && KSTEPS % SPC == 0), "channel split: no skip operand, whole chunks");
};

// mbarriers: X stage full, S skip stage full, R shortcut full, TF accumulators full, TE accumulators drained,
// ST stage consumed by the transform, full[NST] / empty[NST] weight ring
enum { SB_X = 0, SB_S = 1, SB_R = 2, SB_TF = 3, SB_TE = 4, SB_ST = 5, SB_FULL = 6, SB_EMPTY = 9, SB_COUNT = 12 };

__device__ __forceinline__ void worker_barrier() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// stage (TMA box order) -> planes [2C/8][PPAD][8]; TR: the box is [pu][pv][C], planes are indexed pv*PU+pu
template <typename CF, bool TR>
__device__ __forceinline__ void transform_tile(uint32_t stage, uint32_t planes, int tid) {
  constexpr int ITEMS = CF::P * CF::CH;
  constexpr int SH = CF::CH == 8 ? 3 : 4;
#pragma unroll 2
  for (int idx = tid; idx < ITEMS; idx += 256) {
    const int ch = idx & (CF::CH - 1);
    const int sp = idx >> SH;
    int pp = sp;
    if (TR) {
      const int pu = sp / CF::PV, pv = sp - pu * CF::PV;
      pp = pv * CF::PU + pu;
    }
    bf16x8 in, o0, o1;
    in.v = lds128(stage + (uint32_t)idx * 16u);
    concat_elu8(in.v, o0.v, o1.v);
    const uint32_t d0 = planes + (uint32_t)(ch * CF::PPAD + pp) * 16u;
    sts128(d0, o0.v);
    sts128(d0 + (uint32_t)(CF::CH * CF::PPAD * 16), o1.v);
  }
}

// with tf.nn.dropout on the concat_elu output (training): same mask as every other kernel, see conv_s1p.cu
template <typename CF, bool TR>
__device__ __forceinline__ void transform_tile_dropout(uint32_t stage, uint32_t planes, int tid, const Params& p, int b, int y0,
                                                       int x0) {
  constexpr int ITEMS = CF::P * CF::CH;
  constexpr int SH = CF::CH == 8 ? 3 : 4;
  const unsigned long long seed = p.dropout_seed + (p.dropout_seed_dev ? *p.dropout_seed_dev : 0ull);
  for (int idx = tid; idx < ITEMS; idx += 256) {
    const int ch = idx & (CF::CH - 1);
    const int sp = idx >> SH;
    int pu, pv;
    if (TR) { pu = sp / CF::PV; pv = sp - pu * CF::PV; }
    else { pv = sp / CF::PU; pu = sp - pv * CF::PU; }
    const int pp = pv * CF::PU + pu;
    // plane axes (pv, pu) are image (y, x), or (x, y) in the transposed orientation
    const int iy = (TR ? pu : pv) + y0 - 1, ix = (TR ? pv : pu) + x0 - 1;
    const unsigned long long base = (((unsigned long long)b * p.H + iy) * p.W + ix) * (2 * CF::C) + ch * 8;
    const bool inside = iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
    bf16x8 in, o0, o1;
    in.v = lds128(stage + (uint32_t)idx * 16u);
    float sc0[8], sc1[8];
    if (inside) {
      dropout_scale8(seed, base, p.keep_prob, sc0);
      dropout_scale8(seed, base + CF::C, p.keep_prob, sc1);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v = __bfloat162float(in.h[j]);
      const float t = __expf(-fabsf(v)) - 1.f;
      float a0 = __bfloat162float(__float2bfloat16(v > 0.f ? v : t));
      float a1 = __bfloat162float(__float2bfloat16(v > 0.f ? t : -v));
      if (inside) {
        a0 *= sc0[j];
        a1 *= sc1[j];
      }
      o0.h[j] = __float2bfloat16(a0);
      o1.h[j] = __float2bfloat16(a1);
    }
    const uint32_t d0 = planes + (uint32_t)(ch * CF::PPAD + pp) * 16u;
    sts128(d0, o0.v);
    sts128(d0 + (uint32_t)(CF::CH * CF::PPAD * 16), o1.v);
  }
}


template <int CLOG, int J, bool GATE, bool SKIP, bool TR, int NST, int NSPLIT = 1>
__global__ void __launch_bounds__(THREADS) k_conv_s1d(const __grid_constant__ Params p) {
  using CF = Cfg<CLOG, J, GATE, SKIP, TR, NST, NSPLIT>;
  // no static shared memory in these kernels: the dynamic window starts on a 1024-byte boundary, so the base is a
  // compile-time constant of the shared state space (TMA destinations want 128-byte alignment)
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = ptx::smem_u32(smem);
  auto bar = [&](int i) { return sbase + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + 128);
  float* sbias = reinterpret_cast<float*>(smem + 256);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int tid = threadIdx.x;

  // tile -> (image, fast-axis origin u0, slow-axis origin v0); TR: u runs along y
  auto tile_coords = [&](int t, int& b, int& u0, int& v0) {
    const int tu = t % p.tiles_u;
    const int r = t / p.tiles_u;
    const int tv = r % p.tiles_v;
    b = r / p.tiles_v;
    u0 = tu * 8 * J;
    v0 = tv * 16;
  };
  // CTA -> (tile slot, channel half)
  const int nh = (int)blockIdx.x % NSPLIT;
  const int cta = (int)blockIdx.x / NSPLIT, ncta = (int)gridDim.x / NSPLIT;
  const int n_local = (cta < p.ntiles) ? (p.ntiles - 1 - cta) / ncta + 1 : 0;

  if (warp == 8) {
    ptx::tmem_alloc(sbase + 128, (uint32_t)CF::TMEM_COLS);
    ptx::tmem_relinquish();
    if (lane == 0) {
      ptx::mbar_init(bar(SB_X), 1);
      ptx::mbar_init(bar(SB_S), 1);
      ptx::mbar_init(bar(SB_R), 1);
      ptx::mbar_init(bar(SB_TF), (uint32_t)J);
      ptx::mbar_init(bar(SB_TE), 256);
      ptx::mbar_init(bar(SB_ST), 256);
      for (int s = 0; s < NST; ++s) {
        ptx::mbar_init(bar(SB_FULL + s), 1);
        ptx::mbar_init(bar(SB_EMPTY + s), (uint32_t)J);
      }
      ptx::fence_mbar_init();
    }
    __syncwarp();
  }
  pdl_wait();                  // everything above overlapped the previous kernel's tail; global memory from here on
  pdl_launch_dependents();
  for (int i = tid; i < CF::N; i += THREADS) {
    float bv = i < p.Cout ? __ldg(p.bias + i) : 0.f;
    if (GATE && i >= CF::F) bv *= 0.5f;
    sbias[i] = bv;
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  bool ok = true;

  if (warp == 8) {
    // =========================== producer ===========================
    if (lane == 0) {
      // the packed weights are cold in L2 at kernel start (evicted by the activation traffic of the other levels):
      // request the whole image now, the ring's chunk copies then hit L2
      for (int o = 0; o < CF::KSTEPS * CF::N * 32; o += 32768) {
        const int n = CF::KSTEPS * CF::N * 32 - o;
        ptx::prefetch_l2(reinterpret_cast<const unsigned char*>(p.wpk) + o, (uint32_t)(n < 32768 ? n : 32768));
      }
      const uint64_t pol = ptx::policy_evict_last();
      int gchunk = 0;                                   // running weight-chunk counter across tiles
      for (int it = 0; it < n_local && ok; ++it) {
        const uint32_t par_it = (uint32_t)(it & 1), par_prev = par_it ^ 1u;
        int b, u0, v0;
        tile_coords(cta + it * ncta, b, u0, v0);
        const int x0 = TR ? v0 : u0, y0 = TR ? u0 : v0;
        if (it >= 1) {                                  // stage region (aliased by the shortcut tile) free again
          ok = ptx::mbar_wait(bar(SB_TE), par_prev);
          if (!ok) { atomicCAS(p.status, 0, 31); break; }
        }
        ptx::mbar_arrive_expect_tx(bar(SB_X), (uint32_t)(CF::P * CF::C * 2));
        tma_load_4d(sbase + CF::OFF_STAGE, &p.tm_x, bar(SB_X), 0, x0 - 1, y0 - 1, b);
        if (SKIP) {
          ptx::mbar_arrive_expect_tx(bar(SB_S), (uint32_t)(CF::P * CF::C * 2));
          tma_load_4d(sbase + CF::OFF_STAGE_S, &p.tm_skip, bar(SB_S), 0, x0 - 1, y0 - 1, b);
        }
        bool res_issued = !(GATE && p.res_mode == 1);
        for (int c = 0; c < CF::NCHUNKS && ok; ++c, ++gchunk) {
          const int s = gchunk % NST;
          if (gchunk >= NST) {
            ok = ptx::mbar_wait(bar(SB_EMPTY + s), (uint32_t)(((gchunk / NST) - 1) & 1));
            if (!ok) { atomicCAS(p.status, 0, 32); break; }
          }
          if (NSPLIT == 1) {
            const int steps = (c == CF::NCHUNKS - 1) ? (CF::KSTEPS - c * CF::SPC) : CF::SPC;
            const uint32_t bytes = (uint32_t)(steps * CF::N * 32);
            ptx::mbar_arrive_expect_tx(bar(SB_FULL + s), bytes);
            ptx::bulk_g2s_hint(sbase + CF::OFF_RING + s * CHUNK_BYTES,
                               reinterpret_cast<const unsigned char*>(p.wpk) + (size_t)c * CF::SPC * CF::N * 32, bytes, bar(SB_FULL + s), pol);
          } else {
            // one tensor copy per chunk: this CTA's FL rows of the u block (and of the v block) for both K halves of SPC
            // k-steps, rows of FL * 16 contiguous bytes, landing as [kstep][khalf][u rows | v rows][16 B]
            ptx::mbar_arrive_expect_tx(bar(SB_FULL + s), (uint32_t)(CF::SPC * CF::NL * 32));
            tma_load_4d(sbase + CF::OFF_RING + s * CHUNK_BYTES, &p.tm_w, bar(SB_FULL + s), 2 * nh * CF::FL, 0, 0, c * CF::SPC);
          }
          if (!res_issued && c >= NST - 1) {            // ring primed: now fetch the shortcut once the stage is consumed
            ok = ptx::mbar_wait(bar(SB_ST), par_it);
            if (!ok) { atomicCAS(p.status, 0, 33); break; }
            ptx::mbar_arrive_expect_tx(bar(SB_R), (uint32_t)CF::RES_BYTES);
            if (TR) tma_load_4d(sbase + CF::OFF_STAGE, &p.tm_res, bar(SB_R), 0, v0, u0, b);
            else tma_load_4d(sbase + CF::OFF_STAGE, &p.tm_res, bar(SB_R), 0, u0, v0, b);
            res_issued = true;
          }
        }
        if (!res_issued && ok) {
          ok = ptx::mbar_wait(bar(SB_ST), par_it);
          if (!ok) { atomicCAS(p.status, 0, 33); break; }
          ptx::mbar_arrive_expect_tx(bar(SB_R), (uint32_t)CF::RES_BYTES);
          if (TR) tma_load_4d(sbase + CF::OFF_STAGE, &p.tm_res, bar(SB_R), 0, v0, u0, b);
          else tma_load_4d(sbase + CF::OFF_STAGE, &p.tm_res, bar(SB_R), 0, u0, v0, b);
        }
      }
    }
    __syncwarp();
  } else {
    // =========================== workers ===========================
    const uint32_t leader = ptx::elect_one();
    constexpr uint32_t idesc = ptx::idesc_bf16_f32(128, CF::NL);
    const int q = warp & 3, half = warp >> 2;
    const int r = q * 32 + lane;                      // accumulator row == TMEM lane
    const int v = r >> 3, u = r & 7;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    int gchunk = 0;
    for (int it = 0; it < n_local; ++it) {
      const uint32_t par_it = (uint32_t)(it & 1);
      int b, u0, v0;
      tile_coords(cta + it * ncta, b, u0, v0);
      // ---- transform
      ok = ok && ptx::mbar_wait(bar(SB_X), par_it);
      if (!ok) atomicCAS(p.status, 0, 34);
      if (p.keep_prob < 1.0f)
        transform_tile_dropout<CF, TR>(sbase + CF::OFF_STAGE, sbase + CF::OFF_PLANES, tid, p, b, TR ? u0 : v0, TR ? v0 : u0);
      else
        transform_tile<CF, TR>(sbase + CF::OFF_STAGE, sbase + CF::OFF_PLANES, tid);
      if (SKIP) {
        ok = ok && ptx::mbar_wait(bar(SB_S), par_it);
        if (!ok) atomicCAS(p.status, 0, 35);
        transform_tile<CF, TR>(sbase + CF::OFF_STAGE_S, sbase + CF::OFF_PLANES_S, tid);
      }
      ptx::mbar_arrive(bar(SB_ST));                   // stage consumed: the shortcut tile may land there
      ptx::fence_proxy_async_smem();
      worker_barrier();
      // ---- MMA issue (warps 0..J-1): sub-tile j = warp
      if (warp < J) {
        ptx::tc_fence_after_sync();
        const uint64_t a_tmpl = ptx::smem_desc(sbase + CF::OFF_PLANES, CF::PPAD * 16, CF::PU * 16) + (uint64_t)(8 * warp);
        const uint64_t s_tmpl = ptx::smem_desc(sbase + CF::OFF_PLANES_S + (CF::PU + 1) * 16, CF::PPAD * 16, CF::PU * 16) +
                                (uint64_t)(8 * warp);
        const uint64_t b_tmpl = ptx::smem_desc(sbase + CF::OFF_RING, CF::NL * 16, 128);
        const uint32_t d_tmem = tmem_base + (uint32_t)(warp * CF::NL);
#pragma unroll
        for (int c = 0; c < CF::NCHUNKS; ++c) {
          const int gc = gchunk + c;
          const int s = gc % NST;
          ok = ok && ptx::mbar_wait(bar(SB_FULL + s), (uint32_t)((gc / NST) & 1));
          if (!ok) atomicCAS(p.status, 0, 36);
          ptx::tc_fence_after_sync();
          const uint64_t b_chunk = b_tmpl + (uint64_t)((uint32_t)(s * CHUNK_BYTES) >> 4);
#pragma unroll
          for (int i = 0; i < CF::SPC; ++i) {


            const int ks = c * CF::SPC + i;
            if (ks < CF::KSTEPS) {
              const uint32_t b_off = (uint32_t)(i * CF::NL * 32) >> 4;
              if (ks < 9 * CF::KS) {
                const int tap = ks / CF::KS, kk = ks % CF::KS;
                const int di = tap / 3, dj = tap % 3;
                const int dv = TR ? dj : di, du = TR ? di : dj;
                const uint32_t a_off = (uint32_t)((dv * CF::PU + du) * 16 + kk * 2 * CF::PPAD * 16) >> 4;
                if (leader) ptx::mma_bf16_ss(d_tmem, a_tmpl + a_off, b_chunk + b_off, idesc, ks ? 1u : 0u);
              } else {
                const int kk = ks - 9 * CF::KS;
                const uint32_t a_off = (uint32_t)(kk * 2 * CF::PPAD * 16) >> 4;
                if (leader) ptx::mma_bf16_ss(d_tmem, s_tmpl + a_off, b_chunk + b_off, idesc, 1u);
              }
            }
          }
          if (leader) ptx::mma_commit(bar(SB_EMPTY + s));
        }
        if (leader) ptx::mma_commit(bar(SB_TF));
        __syncwarp();
      }
      gchunk += CF::NCHUNKS;
      // ---- epilogue: quadrant q, channel half `half`
      ok = ok && ptx::mbar_wait(bar(SB_TF), par_it);
      if (!ok) atomicCAS(p.status, 0, 37);
      ptx::tc_fence_after_sync();
      if (GATE && p.res_mode == 1) {
        ok = ok && ptx::mbar_wait(bar(SB_R), par_it);
        if (!ok) atomicCAS(p.status, 0, 38);
      }
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const int cu = u0 + 8 * j + u, cv = v0 + v;
        const int oy = TR ? cu : cv, ox = TR ? cv : cu;
        const bool live = ok && oy < p.H && ox < p.W;
        const size_t pix = ((size_t)b * p.H + oy) * p.W + ox;
        const uint32_t col0 = lane_base + (uint32_t)(j * CF::NL);
        // shortcut tile layout = TMA box order: [16 rows of v][8J of u] or, TR, [8J of u][16 of v]
        const uint32_t res_pix = TR ? (uint32_t)((8 * j + u) * 16 + v) : (uint32_t)(v * 8 * J + 8 * j + u);
#pragma unroll 2
        for (int g8 = 0; g8 < CF::FL / 16; ++g8) {
          const int cl = half * (CF::FL / 2) + g8 * 8;      // column inside this CTA's accumulator
          const int c0 = nh * CF::FL + cl;                  // output channel
          float o[8];
          if (GATE) {
            float a[8], g[8];
            ptx::tmem_ld8x2(col0 + cl, col0 + CF::FL + cl, a, g);
            if (p.gate_bwd) {          // du = gy*s, dv = gy*u*s*(1-s), gy = the staged "res" tile
              bf16x8 gy, du, dv;
              gy.v = lds128(sbase + CF::OFF_STAGE + (res_pix * CF::F + (uint32_t)c0) * 2u);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float sg = fmaf(tanh_approx(fmaf(g[i], 0.5f, sbias[CF::F + c0 + i])), 0.5f, 0.5f);
                const float gs = __bfloat162float(gy.h[i]) * sg;
                du.h[i] = __float2bfloat16(gs);
                dv.h[i] = __float2bfloat16(gs * (a[i] + sbias[c0 + i]) * (1.f - sg));
              }
              if (live) {
                *reinterpret_cast<uint4*>(p.y + pix * (2 * CF::F) + c0) = du.v;
                *reinterpret_cast<uint4*>(p.y + pix * (2 * CF::F) + CF::F + c0) = dv.v;
              }
              continue;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float th = tanh_approx(fmaf(g[i], 0.5f, sbias[CF::F + c0 + i]));
              o[i] = (a[i] + sbias[c0 + i]) * fmaf(th, 0.5f, 0.5f);
            }
            if (p.res_mode == 1) {
              bf16x8 rr;
              rr.v = lds128(sbase + CF::OFF_STAGE + (res_pix * CF::F + (uint32_t)c0) * 2u);
#pragma unroll
              for (int i = 0; i < 8; ++i) o[i] += __bfloat162float(rr.h[i]);
            } else if (p.res_mode == 2 && live) {
              shortcut_add8(p.res, p.Cres, p.RH, p.RW, p.res_pool, CF::F, b, oy, ox, c0, o);
            }
          } else {
            ptx::tmem_ld8(col0 + cl, o);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] += sbias[c0 + i];
          }
          if (live) {
            bf16x8 st;
#pragma unroll
            for (int i = 0; i < 8; ++i) st.h[i] = __float2bfloat16(o[i]);
            *reinterpret_cast<uint4*>(p.y + pix * CF::F + c0) = st.v;
          }
        }
      }
      ptx::tc_fence_before_sync();
      ptx::mbar_arrive(bar(SB_TE));                   // accumulators, planes and the stage/shortcut region are free
      if (it + 1 < n_local) {                         // next tile re-uses planes and TMEM: wait for everybody
        worker_barrier();
        ptx::tc_fence_after_sync();
      }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 8) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, (uint32_t)CF::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------ host side

template <int CLOG, int J, bool GATE, bool SKIP, bool TR, int NST, int NSPLIT = 1>
static int launch(const fn_conv_desc& d, int* status, cudaStream_t stream) {
  using CF = Cfg<CLOG, J, GATE, SKIP, TR, NST, NSPLIT>;
  Params p;
  memset(&p, 0, sizeof(p));
  // box order is (C, W-extent, H-extent): the fast tile axis u lies along x unless TR
  const int bw = TR ? CF::PV : CF::PU, bh = TR ? CF::PU : CF::PV;
  if (!make_map_nhwc(&p.tm_x, d.x, d.B, d.H, d.W, CF::C, bw, bh)) { set_error("conv_s1d: tensor map (x) failed"); return FN_E_UNSUPPORTED; }
  if (SKIP && !make_map_nhwc(&p.tm_skip, d.skip, d.B, d.SH, d.SW, CF::C, bw, bh)) {
    set_error("conv_s1d: tensor map (skip) failed");
    return FN_E_UNSUPPORTED;
  }
  p.res_mode = 0;
  if (GATE && d.res != nullptr) {
    if (!d.res_pool && d.Cres == CF::F) {
      p.res_mode = 1;
      if (!make_map_nhwc(&p.tm_res, d.res, d.B, d.RH, d.RW, CF::F, TR ? 16 : 8 * J, TR ? 8 * J : 16)) {
        set_error("conv_s1d: tensor map (res) failed");
        return FN_E_UNSUPPORTED;
      }
    } else {
      p.res_mode = 2;
      p.res = (const __nv_bfloat16*)d.res;
    }
  }
  p.wpk = (const __nv_bfloat16*)d.w_tc;
  p.bias = d.bias;
  p.y = (__nv_bfloat16*)d.y;
  p.status = status;
  p.B = d.B; p.H = d.H; p.W = d.W;
  const int EU = TR ? d.H : d.W, EV = TR ? d.W : d.H;
  p.tiles_u = (EU + 8 * J - 1) / (8 * J);
  p.tiles_v = (EV + 15) / 16;
  p.ntiles = d.B * p.tiles_u * p.tiles_v;
  p.Cout = d.Cout;
  p.Cres = d.Cres; p.RH = d.RH; p.RW = d.RW; p.res_pool = d.res_pool;
  p.keep_prob = d.keep_prob;
  p.dropout_seed = d.dropout_seed;
  p.dropout_seed_dev = (const unsigned long long*)d.dropout_seed_dev;
  p.gate_bwd = d.epilogue == FN_EPI_GATE_BWD ? 1 : 0;
  auto kern = k_conv_s1d<CLOG, J, GATE, SKIP, TR, NST, NSPLIT>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(s1d): %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  if (NSPLIT > 1) {
    // packed weights [kstep][khalf][N rows][8 bf16] seen as 8-byte units: (2 F per block of F rows, N / F blocks, 2, ksteps)
    PFN_encodeTiled enc = get_encode();
    const cuuint64_t dims[4] = {(cuuint64_t)(2 * CF::F), (cuuint64_t)CF::PARTS, 2, (cuuint64_t)CF::KSTEPS};
    const cuuint64_t strides[3] = {(cuuint64_t)CF::F * 16, (cuuint64_t)CF::N * 16, (cuuint64_t)CF::N * 32};
    const cuuint32_t box[4] = {(cuuint32_t)(2 * CF::FL), (cuuint32_t)CF::PARTS, 2, (cuuint32_t)CF::SPC};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    if (!enc || enc(&p.tm_w, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void*>(d.w_tc), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      set_error("conv_s1d: tensor map (weights) failed");
      return FN_E_UNSUPPORTED;
    }
  }
  int grid = (p.ntiles < 148 / NSPLIT ? p.ntiles : 148 / NSPLIT) * NSPLIT;
  launch_pdl(kern, dim3(grid), dim3(THREADS), (size_t)CF::SMEM, stream, p);
  count_launch();
  return check_launch("k_conv_s1d");
}

static int g_nsplit = 1;    // debug: 0 disables the channel split (A/B timing, tests)

}  // namespace s1d

int* tc_status_word();

// Returns true if a deep-layer kernel exists for this launch (and runs it unless dry_run).
bool conv_s1d_try(const fn_conv_desc& d, cudaStream_t stream, bool dry_run, int& rc) {
  using namespace s1d;
  rc = 0;
  if (d.ksize != 3 || d.stride != 1 || d.prologue != FN_PRO_CONCAT_ELU || d.w_tc == nullptr) return false;
  if (d.keep_prob < 1.0f && (d.skip != nullptr || d.dropout_offset != 0)) return false;
  const bool gate = d.epilogue == FN_EPI_GATE_RES || d.epilogue == FN_EPI_GATE_BWD;
  if (!gate && d.epilogue != FN_EPI_BIAS) return false;
  if (d.Cout != (gate ? 2 * d.Cin : d.Cin)) return false;
  if (d.epilogue == FN_EPI_GATE_BWD && (d.res == nullptr || d.res_pool || d.Cres != d.Cin)) return false;
  const bool skip = d.skip != nullptr;
  if (skip && (gate || d.Cskip != d.Cin)) return false;
  const int clog = d.Cin == 64 ? 3 : d.Cin == 128 ? 4 : -1;
  if (clog < 0) return false;
  if (get_encode() == nullptr) return false;
  const bool tr = d.H < 16 && d.W >= 16;        // 8 x 16 images: one transposed tile per image
  if (tr && (d.H % 8) != 0) return false;
  if (!tr && d.H < 16) return false;
  const int EU = tr ? d.H : d.W;
#define SD_CASE(CL, JJ, G, S, T, NS)                                                                    \
  if (clog == CL && gate == G && skip == S && tr == T && ((EU % (8 * JJ)) == 0 || JJ == 1) &&            \
      Cfg<CL, JJ, G, S, T, NS>::SMEM <= 226 * 1024) {                                                    \
    if (!dry_run) {                                                                                      \
      int* st = tc_status_word();                                                                        \
      constexpr int SPL = (CL == 4 && T && !S) ? 2 : 1;   /* 8x16 level: channel split when under half a wave */ \
      const bool split = SPL == 2 && g_nsplit && d.B * ((d.W + 15) / 16) * ((d.H + 8 * JJ - 1) / (8 * JJ)) <= 74;     \
      rc = !st ? (int)cudaErrorMemoryAllocation                                                          \
               : split ? launch<CL, JJ, G, S, T, NS, SPL>(d, st, stream) : launch<CL, JJ, G, S, T, NS>(d, st, stream); \
    }                                                                                                    \
    return true;                                                                                         \
  }
  // C = 64 (16x32 level): two sub-tiles share every weight chunk where shared memory allows
  SD_CASE(3, 2, true, false, false, 3) SD_CASE(3, 2, true, false, false, 2) SD_CASE(3, 1, true, false, false, 3)
  SD_CASE(3, 2, false, false, false, 3) SD_CASE(3, 1, false, false, false, 3)
  SD_CASE(3, 2, false, true, false, 2) SD_CASE(3, 1, false, true, false, 3) SD_CASE(3, 1, false, true, false, 2)
  SD_CASE(3, 1, true, false, true, 3) SD_CASE(3, 1, false, false, true, 3) SD_CASE(3, 1, false, true, true, 2)
  // C = 128 (8x16 level, transposed; also 16x32-or-larger images at wider filter settings)
  SD_CASE(4, 1, true, false, true, 3) SD_CASE(4, 1, true, false, true, 2)
  SD_CASE(4, 1, false, false, true, 3) SD_CASE(4, 1, false, false, true, 2)
  SD_CASE(4, 1, false, true, true, 2)
  SD_CASE(4, 1, true, false, false, 3) SD_CASE(4, 1, true, false, false, 2)
  SD_CASE(4, 1, false, false, false, 3) SD_CASE(4, 1, false, false, false, 2)
  SD_CASE(4, 1, false, true, false, 2)
#undef SD_CASE
  return false;
}

}  // namespace fn

extern "C" {
// test / timing hook: 0 disables the two-CTA channel split of the 8x16 level
void fn_debug_set_nsplit(int v) { fn::s1d::g_nsplit = v; }
}
