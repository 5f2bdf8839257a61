// Persistent, warp-specialised, template-specialised tcgen05 kernel for the stride-1 3x3
// convolutions of flow_net with the concat_elu prologue and resident weights (C in {8,16,32}):
// the 128x256, 64x128 and 32x64 levels.  These launches are HBM- and instruction-issue-bound, so
// everything that can be a compile-time constant is one and the three phases of a tile run
// concurrently on different warps:
//
//   warp 8      producer   TMA tensor loads (cp.async.bulk.tensor.4d) of the raw halo tile, the skip
//                          tile and the shortcut tile, two tiles ahead; out-of-image pixels are
//                          zero-filled by the hardware (= TF 'SAME' padding); weights once per CTA
//   warps 0-3   front      stage -> concat_elu (one ex2 per element) -> UMMA A planes (shared ->
//                          registers -> shared), then one elected lane per sub-tile issues the fully
//                          unrolled tcgen05.mma program (immediate descriptor offsets) into TMEM
//   warps 4-7   back       TMEM -> registers -> bias, gate*sigmoid, shortcut -> bf16 NHWC stores
//
// Stage buffers, A planes, TMEM accumulators and shortcut tiles are double-buffered; mbarriers carry
// the hand-offs (X: TMA->front, XF: front->producer, TF: MMA->back and planes-free->front,
// TE: back->issuers and ->producer, R: shortcut TMA->back).  Same math, A/B layouts and MMA order as
// the generic k_conv_tc (conv_tc.cu); results are bit-reproducible run to run.
// reference semantics: res_block conv_1 (+nin) and conv_2 (+gate, +shortcut), flow_architecture.py:129-165
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"
#include "tma.cuh"

namespace fn {
using namespace tma;

constexpr int SP_THREADS = 448;   // warps 0-3 front, 4-11 back, 12 TMA producer (+TMEM owner), 13 MMA issuer
constexpr int SP_WARP_PRODUCER = 12, SP_WARP_ISSUER = 13;
constexpr int SP_HDR = 256 + 1024;

struct SpParams {
  CUtensorMap tm_x, tm_skip, tm_res;   // 64-byte aligned by the type
  const __nv_bfloat16* wpk;
  const float* bias;
  __nv_bfloat16* y;
  const __nv_bfloat16* res;            // generic shortcut path (pool / channel pad); nullptr if none or staged
  int* status;
  int B, H, W;
  int tiles_u, tiles_v, ntiles;
  int Cout;                            // real accumulator width (bias entries)
  int res_mode;                        // 0 none, 1 staged plain tile (Cres == F), 2 generic global loads
  int Cres, RH, RW, res_pool;
  int gate_bwd;                        // FN_EPI_GATE_BWD: y = d(pre-gate) [., 2F] from the recomputed [u|v] and g = res
  float keep_prob;                     // < 1: dropout on the prologue output (training)
  unsigned long long dropout_seed;
  const unsigned long long* dropout_seed_dev;   // optional device-resident addend (graph replays)
  long long* timing;                   // debug: 8 clock64 stamps per CTA for its third tile (nullptr in production)
  int timing_ctas;
};

template <int CLOG, int J, bool GATE, bool SKIP>
struct SpCfg {
  static constexpr int CH = 1 << CLOG;           // 8-channel chunks of the raw input
  static constexpr int C = 8 * CH;               // raw input channels
  static constexpr int F = C;                    // channels written per pixel
  static constexpr int NREAL = GATE ? 2 * C : C;
  static constexpr int N = NREAL < 16 ? 16 : NREAL;
  static constexpr int KS = CH;                  // k-steps per tap (Keff = 2C)
  static constexpr int KSTEPS = 9 * KS + (SKIP ? KS : 0);
  static constexpr int PU = 8 * J + 2, PV = 18, P = PU * PV;
  // plane stride (pixels) that keeps the 16-byte stores of 8 neighbouring lanes on distinct banks
  static constexpr int PPAD = CH == 1 ? P : (P + ((CH >= 8 ? 1 : (CH == 4 ? 2 : 4)) - (P & 7) + 8) % 8);
  static constexpr int W_BYTES = KSTEPS * N * 32;
  static constexpr int PLANE_BYTES = (2 * CH * PPAD * 16 + 127) / 128 * 128;   // one operand, 2C channels
  static constexpr int STAGE_BYTES = ((P * C * 2) + 127) / 128 * 128;
  static constexpr int RES_BYTES = GATE ? 128 * J * F * 2 : 0;
  static constexpr int OFF_W = SP_HDR;
  static constexpr int OFF_PLANES = OFF_W + (W_BYTES + 127) / 128 * 128;       // [2]
  static constexpr int OFF_PLANES_S = OFF_PLANES + 2 * PLANE_BYTES;            // [2] if SKIP
  static constexpr int OFF_STAGE = OFF_PLANES_S + (SKIP ? 2 * PLANE_BYTES : 0);  // [2]
  static constexpr int OFF_STAGE_S = OFF_STAGE + 2 * STAGE_BYTES;              // [2] if SKIP
  static constexpr int OFF_RES = OFF_STAGE_S + (SKIP ? 2 * STAGE_BYTES : 0);   // [2] if GATE
  static constexpr int SMEM = OFF_RES + 2 * RES_BYTES + 128;                   // +128: base alignment slack
  static constexpr int ACC_COLS = J * N;                                        // one accumulator buffer
  static constexpr int TMEM_COLS = (2 * ACC_COLS <= 32) ? 32 : (2 * ACC_COLS <= 64) ? 64 : (2 * ACC_COLS <= 128) ? 128
                                   : (2 * ACC_COLS <= 256) ? 256 : 512;
};

// mbarriers (8 B each) inside the header; [2] = one per buffer
enum { SB_W = 0, SB_X = 1, SB_S = 3, SB_XF = 5, SB_TF = 7, SB_TE = 9, SB_R = 11, SB_PF = 13, SB_COUNT = 15 };

__device__ __forceinline__ void front_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// stage [pp][C] raw bf16  ->  planes [2C/8][PPAD][8] = [elu(x) | elu(-x)], both bf16 (shared addresses)
template <typename Cfg>
__device__ __forceinline__ void transform_tile(uint32_t stage, uint32_t planes, int tid) {
  constexpr int ITEMS = Cfg::P * Cfg::CH;
  constexpr int SH = Cfg::CH == 1 ? 0 : (Cfg::CH == 2 ? 1 : (Cfg::CH == 4 ? 2 : 3));
#pragma unroll 4
  for (int idx = tid; idx < ITEMS; idx += 128) {
    const int ch = idx & (Cfg::CH - 1);
    const int pp = idx >> SH;
    bf16x8 in, o0, o1;
    in.v = lds128(stage + (uint32_t)idx * 16u);
    concat_elu8(in.v, o0.v, o1.v);
    const uint32_t d0 = planes + (uint32_t)(ch * Cfg::PPAD + pp) * 16u;
    sts128(d0, o0.v);
    sts128(d0 + (uint32_t)(Cfg::CH * Cfg::PPAD * 16), o1.v);
  }
}

// the same with tf.nn.dropout on the concat_elu output (flow_architecture.py:144-145; training only): element k of
// input pixel pidx is scaled by dropout_scale(seed, pidx*2C + k) -- the mask every other kernel of the library uses
template <typename Cfg>
__device__ __forceinline__ void transform_tile_dropout(uint32_t stage, uint32_t planes, int tid, const SpParams& p, int b, int y0,
                                                       int x0) {
  constexpr int ITEMS = Cfg::P * Cfg::CH;
  constexpr int SH = Cfg::CH == 1 ? 0 : (Cfg::CH == 2 ? 1 : (Cfg::CH == 4 ? 2 : 3));
  const unsigned long long seed = p.dropout_seed + (p.dropout_seed_dev ? *p.dropout_seed_dev : 0ull);
  for (int idx = tid; idx < ITEMS; idx += 128) {
    const int ch = idx & (Cfg::CH - 1);
    const int pp = idx >> SH;
    const int pv = pp / Cfg::PU, pu = pp - pv * Cfg::PU;
    const int iy = y0 - 1 + pv, ix = x0 - 1 + pu;
    const unsigned long long base = (((unsigned long long)b * p.H + iy) * p.W + ix) * (2 * Cfg::C) + ch * 8;
    const bool inside = iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
    bf16x8 in, o0, o1;
    in.v = lds128(stage + (uint32_t)idx * 16u);
    float sc0[8], sc1[8];
    if (inside) {
      dropout_scale8(seed, base, p.keep_prob, sc0);
      dropout_scale8(seed, base + Cfg::C, p.keep_prob, sc1);
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
    const uint32_t d0 = planes + (uint32_t)(ch * Cfg::PPAD + pp) * 16u;
    sts128(d0, o0.v);
    sts128(d0 + (uint32_t)(Cfg::CH * Cfg::PPAD * 16), o1.v);
  }
}


template <int CLOG, int J, bool GATE, bool SKIP, bool DROP = false>
__global__ void __launch_bounds__(SP_THREADS) k_conv_s1p(const __grid_constant__ SpParams p) {
  using Cfg = SpCfg<CLOG, J, GATE, SKIP>;
  // no static shared memory in these kernels: the dynamic window starts on a 1024-byte boundary, so the base is a
  // compile-time constant of the shared state space (TMA destinations want 128-byte alignment)
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = ptx::smem_u32(smem);
  auto bar = [&](int i) { return sbase + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + 128);
  float* sbias = reinterpret_cast<float*>(smem + 256);    // [N]; gate: the v half is pre-multiplied by 0.5

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int tid = threadIdx.x;

  auto tile_coords = [&](int t, int& b, int& y0, int& x0) {
    const int tu = t % p.tiles_u;
    const int r = t / p.tiles_u;
    const int tv = r % p.tiles_v;
    b = r / p.tiles_v;
    x0 = tu * 8 * J;
    y0 = tv * 16;
  };
  const int n_local = ((int)blockIdx.x < p.ntiles) ? (p.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  long long* stamps = (p.timing != nullptr && (int)blockIdx.x < p.timing_ctas) ? p.timing + (size_t)blockIdx.x * 8 : nullptr;

  // producer lane: wait for tile it's buffers (it >= 2), then start its tensor copies
  auto issue_tile = [&](int it) -> bool {
    const int s = it & 1;
    const uint32_t prev = (uint32_t)(((it >> 1) - 1) & 1);    // parity of the previous use of buffer s
    int b, y0, x0;
    tile_coords((int)blockIdx.x + it * (int)gridDim.x, b, y0, x0);
    if (it >= 2 && !ptx::mbar_wait(bar(SB_XF + s), prev)) {    // front finished reading stage[s]
      atomicCAS(p.status, 0, 21);
      return false;
    }
    ptx::mbar_arrive_expect_tx(bar(SB_X + s), (uint32_t)(Cfg::P * Cfg::C * 2));
    tma_load_4d(sbase + Cfg::OFF_STAGE + s * Cfg::STAGE_BYTES, &p.tm_x, bar(SB_X + s), 0, x0 - 1, y0 - 1, b);
    if (SKIP) {
      ptx::mbar_arrive_expect_tx(bar(SB_S + s), (uint32_t)(Cfg::P * Cfg::C * 2));
      tma_load_4d(sbase + Cfg::OFF_STAGE_S + s * Cfg::STAGE_BYTES, &p.tm_skip, bar(SB_S + s), 0, x0 - 1, y0 - 1, b);
    }
    if (GATE && p.res_mode == 1) {
      if (it >= 2 && !ptx::mbar_wait(bar(SB_TE + s), prev)) {  // back finished with res[s]
        atomicCAS(p.status, 0, 22);
        return false;
      }
      ptx::mbar_arrive_expect_tx(bar(SB_R + s), (uint32_t)Cfg::RES_BYTES);
      tma_load_4d(sbase + Cfg::OFF_RES + s * Cfg::RES_BYTES, &p.tm_res, bar(SB_R + s), 0, x0, y0, b);
    }
    return true;
  };

  if (threadIdx.x == 64) {       // descriptor fetch off the first tile's critical path
    prefetch_tensormap(&p.tm_x);
    if (SKIP) prefetch_tensormap(&p.tm_skip);
  }
  // ---- one-time setup
  if (warp == SP_WARP_PRODUCER) {
    ptx::tmem_alloc(sbase + 128, (uint32_t)Cfg::TMEM_COLS);
    ptx::tmem_relinquish();
    if (lane == 0) {
      ptx::mbar_init(bar(SB_W), 1);
      for (int s = 0; s < 2; ++s) {
        ptx::mbar_init(bar(SB_X + s), 1);
        ptx::mbar_init(bar(SB_S + s), 1);
        ptx::mbar_init(bar(SB_XF + s), 128);
        ptx::mbar_init(bar(SB_PF + s), 128);
        ptx::mbar_init(bar(SB_TF + s), 1);
        ptx::mbar_init(bar(SB_TE + s), 256);
        ptx::mbar_init(bar(SB_R + s), 1);
      }
      ptx::fence_mbar_init();
      // the weights and the first two tiles need no hand-shake: start them now, under the rest of the set-up
      // (bias staging, TMEM address exchange, the block barrier), instead of after it
      pdl_wait();
      ptx::mbar_arrive_expect_tx(bar(SB_W), (uint32_t)Cfg::W_BYTES);
      ptx::bulk_g2s(sbase + Cfg::OFF_W, p.wpk, (uint32_t)Cfg::W_BYTES, bar(SB_W));
      for (int it = 0; it < 2 && it < n_local; ++it) issue_tile(it);
    }
    __syncwarp();
  }
  pdl_wait();                  // everything above overlapped the previous kernel's tail; global memory from here on
  pdl_launch_dependents();
  for (int i = tid; i < Cfg::N; i += SP_THREADS) {
    float bv = i < p.Cout ? __ldg(p.bias + i) : 0.f;
    if (GATE && i >= Cfg::F) bv *= 0.5f;
    sbias[i] = bv;
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  bool ok = true;

  if (warp == SP_WARP_PRODUCER) {
    // =========================== producer: TMA ===========================
    if (lane == 0) {
      for (int it = 2; it < n_local && ok; ++it) ok = issue_tile(it);     // tiles 0 and 1 went out during set-up
    }
    __syncwarp();
  } else if (warp == SP_WARP_ISSUER) {
    // =========================== issuer: one elected lane feeds the tensor core ===========================
    // The warp runs the loop convergently so descriptor arithmetic stays on the uniform datapath; the whole
    // k-step program of a tile (J sub-tiles x 9 taps x KS) is unrolled with immediate descriptor offsets.
    const uint32_t leader = ptx::elect_one();
    constexpr uint32_t idesc = ptx::idesc_bf16_f32(128, Cfg::N);
    const uint64_t b_tmpl = ptx::smem_desc(sbase + Cfg::OFF_W, Cfg::N * 16, 128);
    ok = ptx::mbar_wait(bar(SB_W), 0);
    if (!ok) atomicCAS(p.status, 0, 13);
    for (int it = 0; it < n_local; ++it) {
      const int s = it & 1;
      const uint32_t cur = (uint32_t)((it >> 1) & 1), prev = cur ^ 1u;
      ok = ok && ptx::mbar_wait(bar(SB_PF + s), cur);                    // planes[s] written and fenced
      if (it >= 2) ok = ok && ptx::mbar_wait(bar(SB_TE + s), prev);      // back drained accumulator buffer s
      if (!ok) atomicCAS(p.status, 0, 16);
      ptx::tc_fence_after_sync();
      if (stamps && it == 2 && lane == 0) stamps[3] = clock64();
      const uint64_t a_tmpl = ptx::smem_desc(sbase + Cfg::OFF_PLANES + s * Cfg::PLANE_BYTES, Cfg::PPAD * 16, Cfg::PU * 16);
      const uint64_t s_tmpl = ptx::smem_desc(sbase + Cfg::OFF_PLANES_S + s * Cfg::PLANE_BYTES + (Cfg::PU + 1) * 16, Cfg::PPAD * 16,
                                             Cfg::PU * 16);
      const uint32_t d_tmem = tmem_base + (uint32_t)(s * Cfg::ACC_COLS);
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
        for (int kk = 0; kk < Cfg::KS; ++kk) {
          const uint32_t a_off = (uint32_t)(((tap / 3) * Cfg::PU + (tap % 3)) * 16 + kk * 2 * Cfg::PPAD * 16) >> 4;
          const uint32_t b_off = (uint32_t)((tap * Cfg::KS + kk) * Cfg::N * 32) >> 4;
#pragma unroll
          for (int j = 0; j < J; ++j)
            if (leader) ptx::mma_bf16_ss(d_tmem + (uint32_t)(j * Cfg::N), a_tmpl + a_off + (uint32_t)(8 * j), b_tmpl + b_off, idesc,
                                         (tap | kk) ? 1u : 0u);
        }
      }
      if (SKIP) {
#pragma unroll
        for (int kk = 0; kk < Cfg::KS; ++kk) {
          const uint32_t a_off = (uint32_t)(kk * 2 * Cfg::PPAD * 16) >> 4;
          const uint32_t b_off = (uint32_t)((9 * Cfg::KS + kk) * Cfg::N * 32) >> 4;
#pragma unroll
          for (int j = 0; j < J; ++j)
            if (leader) ptx::mma_bf16_ss(d_tmem + (uint32_t)(j * Cfg::N), s_tmpl + a_off + (uint32_t)(8 * j), b_tmpl + b_off, idesc, 1u);
        }
      }
      if (leader) ptx::mma_commit(bar(SB_TF + s));
      __syncwarp();
      if (stamps && it == 2 && lane == 0) stamps[4] = clock64();
    }
  } else if (warp < 4) {
    // =========================== front: stage -> concat_elu -> A planes ===========================
    for (int it = 0; it < n_local; ++it) {
      const int s = it & 1;
      const uint32_t cur = (uint32_t)((it >> 1) & 1), prev = cur ^ 1u;
      const uint32_t planes = sbase + Cfg::OFF_PLANES + s * Cfg::PLANE_BYTES;
      const uint32_t planes_s = sbase + Cfg::OFF_PLANES_S + s * Cfg::PLANE_BYTES;
      long long* stamp = (stamps && it == 2 && tid == 0) ? stamps : nullptr;
      if (stamp) stamp[0] = clock64();
      ok = ok && ptx::mbar_wait(bar(SB_X + s), cur);                // raw tile landed
      if (it >= 2) ok = ok && ptx::mbar_wait(bar(SB_TF + s), prev); // MMAs of tile it-2 no longer read planes[s]
      if (!ok) atomicCAS(p.status, 0, 11);
      if (stamp) stamp[1] = clock64();
      if (DROP) {          // a separate instantiation: the mask hashing must not cost the inference kernels registers
        int b, y0, x0;
        tile_coords((int)blockIdx.x + it * (int)gridDim.x, b, y0, x0);
        transform_tile_dropout<Cfg>(sbase + Cfg::OFF_STAGE + s * Cfg::STAGE_BYTES, planes, tid, p, b, y0, x0);
      } else {
        transform_tile<Cfg>(sbase + Cfg::OFF_STAGE + s * Cfg::STAGE_BYTES, planes, tid);
      }
      if (SKIP) {
        ok = ok && ptx::mbar_wait(bar(SB_S + s), cur);
        if (!ok) atomicCAS(p.status, 0, 12);
        transform_tile<Cfg>(sbase + Cfg::OFF_STAGE_S + s * Cfg::STAGE_BYTES, planes_s, tid);
      }
      ptx::mbar_arrive(bar(SB_XF + s));                             // stage[s] may be refilled
      ptx::fence_proxy_async_smem();                                // this thread's plane writes -> tensor-core proxy
      ptx::mbar_arrive(bar(SB_PF + s));                             // planes[s] ready (128 arrivals)
      if (stamp) stamp[2] = clock64();
    }
  } else {
    // =========================== back: 8 epilogue warps ===========================
    // warp wb: TMEM lane quadrant wb & 3; half = wb >> 2 takes the odd / even sub-tiles (J >= 2) or the
    // upper / lower half of the channels (J == 1)
    const int wb = warp - 4;
    const int q = wb & 3, half = wb >> 2;
    const int r = q * 32 + lane;             // accumulator row == TMEM lane
    const int v = r >> 3, u = r & 7;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    constexpr int JSTEP = J >= 2 ? 2 : 1;
    constexpr int CSPAN = J >= 2 ? Cfg::F : (Cfg::F >= 16 ? Cfg::F / 2 : Cfg::F);
    const int j_first = J >= 2 ? half : 0;
    const int c_first = J >= 2 ? 0 : (Cfg::F >= 16 ? half * (Cfg::F / 2) : 0);
    const bool idle = (J == 1 && Cfg::F < 16 && half == 1);   // nothing to split: second half only signals
    for (int it = 0; it < n_local; ++it) {
      const int s = it & 1;
      const uint32_t cur = (uint32_t)((it >> 1) & 1);
      int b, y0, x0;
      tile_coords((int)blockIdx.x + it * (int)gridDim.x, b, y0, x0);
      // 1-channel shortcut (the boundary image under the first block): one scalar per pixel lands in the
      // LAST channel (front zero-pad, reference :158-163); fetch it while the MMAs run
      float r1[J];
#pragma unroll
      for (int j = 0; j < J; ++j) {
        r1[j] = 0.f;
        if (GATE && p.res_mode == 3) {
          const int oy1 = y0 + v, ox1 = x0 + 8 * j + u;
          if (oy1 < p.H && ox1 < p.W) r1[j] = __bfloat162float(__ldg(p.res + ((size_t)b * p.H + oy1) * p.W + ox1));
        }
      }
      long long* bstamp = (stamps && it == 2 && tid == 128) ? stamps : nullptr;
      if (bstamp) bstamp[5] = clock64();
      ok = ok && ptx::mbar_wait(bar(SB_TF + s), cur);
      if (!ok) atomicCAS(p.status, 0, 14);
      ptx::tc_fence_after_sync();
      if (GATE && p.res_mode == 1) {
        ok = ok && ptx::mbar_wait(bar(SB_R + s), cur);
        if (!ok) atomicCAS(p.status, 0, 15);
      }
      if (bstamp) bstamp[6] = clock64();
      const int oy = y0 + v;
      if (!idle) {
#pragma unroll
        for (int jj = 0; jj < J; jj += JSTEP) {
          const int j = jj + j_first;
          if (j >= J) break;
          const int ox = x0 + 8 * j + u;
          const bool live = ok && oy < p.H && ox < p.W;
          const size_t pix = ((size_t)b * p.H + oy) * p.W + ox;
          const uint32_t col0 = lane_base + (uint32_t)(s * Cfg::ACC_COLS + j * Cfg::N);
#pragma unroll
          for (int cc = 0; cc < CSPAN; cc += 8) {
            const int c0 = c_first + cc;
            float o[8];
            if (GATE) {
              float a[8], g[8];
              ptx::tmem_ld8x2(col0 + c0, col0 + Cfg::F + c0, a, g);
              if (p.gate_bwd) {
                // adjoint of y = u*sigmoid(v): du = gy*s, dv = gy*u*s*(1-s); gy is the staged "res" tile
                bf16x8 gy, du, dv;
                gy.v = lds128(sbase + Cfg::OFF_RES + s * Cfg::RES_BYTES + (uint32_t)((v * 8 * J + 8 * j + u) * Cfg::F + c0) * 2u);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float sg = fmaf(tanh_approx(fmaf(g[i], 0.5f, sbias[Cfg::F + c0 + i])), 0.5f, 0.5f);
                  const float gs = __bfloat162float(gy.h[i]) * sg;
                  du.h[i] = __float2bfloat16(gs);
                  dv.h[i] = __float2bfloat16(gs * (a[i] + sbias[c0 + i]) * (1.f - sg));
                }
                if (live) {
                  *reinterpret_cast<uint4*>(p.y + pix * (2 * Cfg::F) + c0) = du.v;
                  *reinterpret_cast<uint4*>(p.y + pix * (2 * Cfg::F) + Cfg::F + c0) = dv.v;
                }
                continue;
              }
              // u * sigmoid(v) with sigmoid(z) = 0.5 + 0.5 tanh(z/2); the v bias is stored pre-halved
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float th = tanh_approx(fmaf(g[i], 0.5f, sbias[Cfg::F + c0 + i]));
                o[i] = (a[i] + sbias[c0 + i]) * fmaf(th, 0.5f, 0.5f);
              }
              if (p.res_mode == 1) {
                bf16x8 rr;
                rr.v = lds128(sbase + Cfg::OFF_RES + s * Cfg::RES_BYTES + (uint32_t)((v * 8 * J + 8 * j + u) * Cfg::F + c0) * 2u);
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] += __bfloat162float(rr.h[i]);
              } else if (p.res_mode == 2 && live) {
                shortcut_add8(p.res, p.Cres, p.RH, p.RW, p.res_pool, Cfg::F, b, oy, ox, c0, o);
              } else if (p.res_mode == 3 && c0 == Cfg::F - 8) {
                o[7] += r1[j];
              }
            } else {
              ptx::tmem_ld8(col0 + c0, o);
#pragma unroll
              for (int i = 0; i < 8; ++i) o[i] += sbias[c0 + i];
            }
            if (live) {
              bf16x8 st;
#pragma unroll
              for (int i = 0; i < 8; ++i) st.h[i] = __float2bfloat16(o[i]);
              *reinterpret_cast<uint4*>(p.y + pix * Cfg::F + c0) = st.v;
            }
          }
        }
      }
      if (bstamp) bstamp[7] = clock64();
      ptx::tc_fence_before_sync();
      ptx::mbar_arrive(bar(SB_TE + s));        // accumulator buffer s and res[s] are free (256 arrivals)
    }
  }

  // ---- teardown
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == SP_WARP_PRODUCER) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, (uint32_t)Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------ host side



int* tc_status_word();   // conv_tc.cu: lazily allocated device status word
void tc_timing_buffer(long long** buf, int* ctas);   // conv_tc.cu: debug stamp buffer (fn_debug_set_timing)

template <int CLOG, int J, bool GATE, bool SKIP, bool DROP = false>
static int launch_s1p(const fn_conv_desc& d, cudaStream_t stream) {
  using Cfg = SpCfg<CLOG, J, GATE, SKIP>;
  SpParams p;
  memset(&p, 0, sizeof(p));
  if (!make_map_nhwc(&p.tm_x, d.x, d.B, d.H, d.W, Cfg::C, Cfg::PU, Cfg::PV)) {
    set_error("conv_s1p: cuTensorMapEncodeTiled(x) failed");
    return FN_E_UNSUPPORTED;
  }
  if (SKIP && !make_map_nhwc(&p.tm_skip, d.skip, d.B, d.SH, d.SW, Cfg::C, Cfg::PU, Cfg::PV)) {
    set_error("conv_s1p: cuTensorMapEncodeTiled(skip) failed");
    return FN_E_UNSUPPORTED;
  }
  p.res_mode = 0;
  if (GATE && d.res != nullptr) {
    if (!d.res_pool && d.Cres == Cfg::F) {
      p.res_mode = 1;
      if (!make_map_nhwc(&p.tm_res, d.res, d.B, d.RH, d.RW, Cfg::F, 8 * J, 16)) {
        set_error("conv_s1p: cuTensorMapEncodeTiled(res) failed");
        return FN_E_UNSUPPORTED;
      }
    } else {
      p.res_mode = (!d.res_pool && d.Cres == 1 && d.RH == d.H && d.RW == d.W) ? 3 : 2;
      p.res = (const __nv_bfloat16*)d.res;
    }
  }
  p.wpk = (const __nv_bfloat16*)d.w_tc;
  p.bias = d.bias;
  p.y = (__nv_bfloat16*)d.y;
  p.status = tc_status_word();
  if (p.status == nullptr) return (int)cudaErrorMemoryAllocation;
  p.B = d.B; p.H = d.H; p.W = d.W;
  p.tiles_u = (d.W + 8 * J - 1) / (8 * J);
  p.tiles_v = (d.H + 15) / 16;
  p.ntiles = d.B * p.tiles_u * p.tiles_v;
  p.Cout = d.Cout;
  p.Cres = d.Cres; p.RH = d.RH; p.RW = d.RW; p.res_pool = d.res_pool;
  p.keep_prob = d.keep_prob;
  p.dropout_seed = d.dropout_seed;
  p.dropout_seed_dev = (const unsigned long long*)d.dropout_seed_dev;
  p.gate_bwd = d.epilogue == FN_EPI_GATE_BWD ? 1 : 0;
  tc_timing_buffer(&p.timing, &p.timing_ctas);

  auto kern = k_conv_s1p<CLOG, J, GATE, SKIP, DROP>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(s1p): %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  int per_sm = (227 * 1024) / (Cfg::SMEM + 1024);
  const int tmem_lim = 512 / Cfg::TMEM_COLS;
  if (per_sm > tmem_lim) per_sm = tmem_lim;
  if (per_sm > 4) per_sm = 4;
  if (per_sm < 1) per_sm = 1;
  int grid = 148 * per_sm;
  if (grid > p.ntiles) grid = p.ntiles;
  launch_pdl(kern, dim3(grid), dim3(SP_THREADS), (size_t)Cfg::SMEM, stream, p);
  count_launch();
  return check_launch("k_conv_s1p");
}

// Returns true if a specialised kernel exists for this launch (and runs it: rc = its return code).
bool conv_s1p_try(const fn_conv_desc& d, cudaStream_t stream, bool dry_run, int& rc) {
  rc = 0;
  if (d.ksize != 3 || d.stride != 1 || d.prologue != FN_PRO_CONCAT_ELU || d.w_tc == nullptr) return false;
  if (d.keep_prob < 1.0f && (d.skip != nullptr || d.dropout_offset != 0)) return false;   // the dropped operand is the conv input
  const bool gate = d.epilogue == FN_EPI_GATE_RES || d.epilogue == FN_EPI_GATE_BWD;
  if (!gate && d.epilogue != FN_EPI_BIAS) return false;
  if (d.Cout != (gate ? 2 * d.Cin : d.Cin)) return false;
  if (d.epilogue == FN_EPI_GATE_BWD && (d.res == nullptr || d.res_pool || d.Cres != d.Cin)) return false;
  const bool skip = d.skip != nullptr;
  if (skip && (gate || d.Cskip != d.Cin)) return false;
  if (d.H < 16) return false;                       // short images use the transposed orientation of k_conv_tc
  if (get_encode() == nullptr) return false;
  const int clog = d.Cin == 8 ? 0 : d.Cin == 16 ? 1 : d.Cin == 32 ? 2 : -1;
  if (clog < 0) return false;
  static const int smem_cap_kb = getenv("FLOWNET_S1P_SMEM_KB") ? atoi(getenv("FLOWNET_S1P_SMEM_KB")) : 112;
  // widest tile that divides the row and still leaves room for 2 CTAs per SM (cases are ordered by J);
  // C = 32 carries up to 73 KB of weights and runs one CTA per SM
#define SP_CASE(CL, JJ, G, S)                                                                     \
  if (clog == CL && ((d.W % (8 * JJ)) == 0 || JJ == 1) && gate == G && skip == S &&               \
      SpCfg<CL, JJ, G, S>::SMEM <= (CL == 2 ? 200 : smem_cap_kb) * 1024) {                        \
    if (!dry_run) rc = (!S && d.keep_prob < 1.0f) ? launch_s1p<CL, JJ, G, false, true>(d, stream)   \
                                                  : launch_s1p<CL, JJ, G, S>(d, stream);            \
    return true;                                                                                  \
  }
  SP_CASE(0, 4, true, false) SP_CASE(0, 4, false, false) SP_CASE(0, 4, false, true)
  SP_CASE(0, 2, true, false) SP_CASE(0, 2, false, false) SP_CASE(0, 2, false, true)
  SP_CASE(0, 1, true, false) SP_CASE(0, 1, false, false) SP_CASE(0, 1, false, true)
  SP_CASE(1, 4, true, false) SP_CASE(1, 4, false, false) SP_CASE(1, 4, false, true)
  SP_CASE(1, 2, true, false) SP_CASE(1, 2, false, false) SP_CASE(1, 2, false, true)
  SP_CASE(1, 1, true, false) SP_CASE(1, 1, false, false) SP_CASE(1, 1, false, true)
  SP_CASE(2, 2, true, false) SP_CASE(2, 2, false, false) SP_CASE(2, 2, false, true)
  SP_CASE(2, 1, true, false) SP_CASE(2, 1, false, false) SP_CASE(2, 1, false, true)
#undef SP_CASE
  return false;
}

}  // namespace fn
