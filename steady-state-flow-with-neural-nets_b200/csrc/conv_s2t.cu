// Persistent, warp-specialised tcgen05 kernels for the two RESAMPLING convolutions of flow_net with
// resident weights, built like k_conv_s1p (conv_s1p.cu: TMA producer warp, 4 front warps, 8 back
// warps, 1 issuer warp, double-buffered stage / planes / TMEM behind mbarriers):
//
//   MODE_S2     3x3 stride-2 'SAME' conv with the concat_elu prologue (conv_1 of the down-sampling
//               res_blocks, reference flow_architecture.py:132-135 with stride=2; TF pads (0,1) on even
//               sizes).  The 33 x (16J+1) input tile is split by the front warps into its four
//               (row, column)-parity sub-planes, so that a stride-2 tap is again a unit-stride view:
//               input 2o+d -> parity d&1, position o + (d>>1).
//   MODE_TCONV  3x3 stride-2 transposed conv (reference :70-87) as a sub-pixel implicit GEMM over INPUT
//               pixels: out[2m+p] takes taps d=p (in[m]) and, for p=0, d=2 (in[m-1]);
//               the four output phases are four TMEM accumulators (4/2/2/1 taps), written by the
//               epilogue to the four interleaved output pixels.  No prologue: the front warps only
//               re-lay the tile out as 8-channel planes.
//
// Same A/B operand layouts, MMA order and bf16 rounding points as the generic k_conv_tc.
#include "common.cuh"
#include "ptx.cuh"
#include "tma.cuh"

namespace fn {
using namespace tma;
namespace s2t {

constexpr int THREADS = 448;
constexpr int WARP_PRODUCER = 12, WARP_ISSUER = 13;
constexpr int HDR = 256 + 1024;
enum { MODE_S2 = 1, MODE_TCONV = 2, MODE_S1N = 3, MODE_S2N = 4 };

struct Params {
  CUtensorMap tm_x;
  const __nv_bfloat16* wpk;
  const float* bias;
  __nv_bfloat16* y;       // raw NHWC output (null: planes only)
  __nv_bfloat16* yp;      // optional planes form of the output (fn_conv_desc.y_planes), plain bias epilogue only
  int* status;
  int B, H, W;            // input tensor
  int OH, OW;             // output tensor
  int RH, RW;             // GEMM-row image: output pixels (S2) or input pixels (TCONV)
  int tiles_u, tiles_v, ntiles;
  int Cout;               // real output channels
  // FN_EPI_PRO_BWD (MODE_S1N): the accumulator is the gradient of a prologue output; y (+)= its adjoint w.r.t. res
  int adj;                // 0: plain bias epilogue; else 1 + enum fn_prologue of the adjoint
  int adj_accumulate;
  const __nv_bfloat16* res;
  float keep_prob;
  unsigned long long seed;
  const unsigned long long* seed_dev;
};

template <int MODE, int CLOG, int J, int NBUF, int R = 1>
struct Cfg {
  static constexpr int CH = 1 << CLOG;                       // 8-channel chunks of the raw input
  static constexpr int C = 8 * CH;
  static constexpr bool S2 = MODE == MODE_S2 || MODE == MODE_S2N;
  static constexpr bool S2RAW = MODE == MODE_S2N;            // stride 2, no prologue (data gradient of a transposed conv)
  static constexpr bool S1N = MODE == MODE_S1N;              // stride 1, no prologue, Cout = R*C (data gradients)
  static constexpr int KPLANES = (S2 && !S2RAW) ? 2 * CH : CH;   // 8-channel planes of one (sub-)operand
  // k-steps per tap; a lone plane (C = 8) is read for both K halves (LBO = 0) against zero-padded weights
  static constexpr bool HALF_K = (!S2 || S2RAW) && CH == 1;
  static constexpr int KS = (S2 && !S2RAW) ? CH : (CH >= 2 ? CH / 2 : 1);
  static constexpr int NREAL = S2 ? 2 * C : (S1N ? R * C : R * C / 2);   // Cout (TCONV, R = 2: the data gradient of a stride-2 conv)
  static constexpr int N = NREAL < 16 ? 16 : NREAL;
  static constexpr int NPH = MODE == MODE_TCONV ? 4 : 1;     // accumulators per sub-tile
  static constexpr int KSTEPS = 9 * KS;
  // stage tile (TMA box): S2 33 x (16J+1) input pixels, TCONV 17 x (8J+1), S1N 18 x (8J+2)
  static constexpr int SW_ = S2 ? 16 * J + 1 : (S1N ? 8 * J + 2 : 8 * J + 1), SH_ = S2 ? 33 : (S1N ? 18 : 17), SP = SW_ * SH_;
  // (sub-)plane geometry
  static constexpr int PU = S1N ? 8 * J + 2 : 8 * J + 1, PV = S1N ? 18 : 17, P = PU * PV;
  static constexpr int PADW = KPLANES >= 8 ? 1 : (KPLANES == 4 ? 2 : (KPLANES == 2 ? 4 : 0));
  static constexpr int PPAD = KPLANES == 1 ? P : P + ((PADW - (P & 7)) + 8) % 8;
  static constexpr int SUB_BYTES = KPLANES * PPAD * 16;
  static constexpr int NSUB = S2 ? 4 : 1;
  static constexpr int W_BYTES = KSTEPS * N * 32;
  static constexpr int PLANE_BYTES = (NSUB * SUB_BYTES + 127) / 128 * 128;
  static constexpr int STAGE_BYTES = (SP * C * 2 + 127) / 128 * 128;
  static constexpr int OFF_W = HDR;
  static constexpr int OFF_PLANES = OFF_W + (W_BYTES + 127) / 128 * 128;
  static constexpr int OFF_STAGE = OFF_PLANES + NBUF * PLANE_BYTES;
  static constexpr int SMEM = OFF_STAGE + NBUF * STAGE_BYTES + 128;
  static constexpr int ACC_COLS = J * NPH * N;
  static constexpr int TCOLS = NBUF * ACC_COLS;
  static constexpr int TMEM_COLS = TCOLS <= 32 ? 32 : TCOLS <= 64 ? 64 : TCOLS <= 128 ? 128 : TCOLS <= 256 ? 256 : 512;
  static_assert(TCOLS <= 512, "accumulators exceed TMEM");
  static_assert(MODE != MODE_TCONV || CH >= 2, "transposed conv needs C >= 16");
};

enum { SB_W = 0, SB_X = 1, SB_XF = 3, SB_PF = 5, SB_TF = 7, SB_TE = 9, SB_COUNT = 11 };


// stage [SH_][SW_][C] raw bf16 -> operand planes
template <typename CF>
__device__ __forceinline__ void transform_tile(uint32_t stage, uint32_t planes, int tid) {
  constexpr int ITEMS = CF::SP * CF::CH;
  constexpr int SH = CF::CH == 1 ? 0 : (CF::CH == 2 ? 1 : (CF::CH == 4 ? 2 : 3));
#pragma unroll 2
  for (int idx = tid; idx < ITEMS; idx += 128) {
    const int ch = idx & (CF::CH - 1);
    const int sp = idx >> SH;                        // stage pixel
    bf16x8 in;
    in.v = lds128(stage + (uint32_t)idx * 16u);
    if (CF::S2) {
      const int r = sp / CF::SW_, c = sp - r * CF::SW_;
      const int sub = (r & 1) * 2 + (c & 1);
      const int pp = (r >> 1) * CF::PU + (c >> 1);
      if (CF::S2RAW) {
        sts128(planes + (uint32_t)(sub * CF::SUB_BYTES) + (uint32_t)(ch * CF::PPAD + pp) * 16u, in.v);
      } else {
        bf16x8 o0, o1;
        concat_elu8(in.v, o0.v, o1.v);
        const uint32_t d0 = planes + (uint32_t)(sub * CF::SUB_BYTES) + (uint32_t)(ch * CF::PPAD + pp) * 16u;
        sts128(d0, o0.v);
        sts128(d0 + (uint32_t)(CF::CH * CF::PPAD * 16), o1.v);
      }
    } else {
      sts128(planes + (uint32_t)(ch * CF::PPAD + sp) * 16u, in.v);
    }
  }
}

template <int MODE, int CLOG, int J, int NBUF, int R = 1>
__global__ void __launch_bounds__(THREADS) k_conv_s2t(const __grid_constant__ Params p) {
  using CF = Cfg<MODE, CLOG, J, NBUF, R>;
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

  auto tile_coords = [&](int t, int& b, int& y0, int& x0) {
    const int tu = t % p.tiles_u;
    const int r = t / p.tiles_u;
    const int tv = r % p.tiles_v;
    b = r / p.tiles_v;
    x0 = tu * 8 * J;
    y0 = tv * 16;
  };
  const int n_local = ((int)blockIdx.x < p.ntiles) ? (p.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  // buffer index and the parity of its k-th use
  auto buf_of = [](int it) { return NBUF == 2 ? (it & 1) : 0; };
  auto use_of = [](int it) { return NBUF == 2 ? (it >> 1) : it; };

  if (warp == WARP_PRODUCER) {
    ptx::tmem_alloc(sbase + 128, (uint32_t)CF::TMEM_COLS);
    ptx::tmem_relinquish();
    if (lane == 0) {
      ptx::mbar_init(bar(SB_W), 1);
      for (int s = 0; s < 2; ++s) {
        ptx::mbar_init(bar(SB_X + s), 1);
        ptx::mbar_init(bar(SB_XF + s), 128);
        ptx::mbar_init(bar(SB_PF + s), 128);
        ptx::mbar_init(bar(SB_TF + s), 1);
        ptx::mbar_init(bar(SB_TE + s), 256);
      }
      ptx::fence_mbar_init();
    }
    __syncwarp();
  }
  pdl_wait();                  // everything above overlapped the previous kernel's tail; global memory from here on
  pdl_launch_dependents();
  for (int i = tid; i < CF::N; i += THREADS) sbias[i] = i < p.Cout ? __ldg(p.bias + i) : 0.f;
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  bool ok = true;

  if (warp == WARP_PRODUCER) {
    // =========================== producer: TMA ===========================
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(bar(SB_W), (uint32_t)CF::W_BYTES);
      ptx::bulk_g2s(sbase + CF::OFF_W, p.wpk, (uint32_t)CF::W_BYTES, bar(SB_W));
      for (int it = 0; it < n_local && ok; ++it) {
        const int s = buf_of(it), k = use_of(it);
        int b, y0, x0;
        tile_coords((int)blockIdx.x + it * (int)gridDim.x, b, y0, x0);
        if (k >= 1) {
          ok = ptx::mbar_wait(bar(SB_XF + s), (uint32_t)((k - 1) & 1));       // front finished reading stage[s]
          if (!ok) { atomicCAS(p.status, 0, 41); break; }
        }
        ptx::mbar_arrive_expect_tx(bar(SB_X + s), (uint32_t)(CF::SP * CF::C * 2));
        if (CF::S2) tma_load_4d(sbase + CF::OFF_STAGE + s * CF::STAGE_BYTES, &p.tm_x, bar(SB_X + s), 0, 2 * x0, 2 * y0, b);
        else tma_load_4d(sbase + CF::OFF_STAGE + s * CF::STAGE_BYTES, &p.tm_x, bar(SB_X + s), 0, x0 - 1, y0 - 1, b);
      }
    }
    __syncwarp();
  } else if (warp == WARP_ISSUER) {
    // =========================== issuer ===========================
    const uint32_t leader = ptx::elect_one();
    constexpr uint32_t idesc = ptx::idesc_bf16_f32(128, CF::N);
    const uint64_t b_tmpl = ptx::smem_desc(sbase + CF::OFF_W, CF::N * 16, 128);
    ok = ptx::mbar_wait(bar(SB_W), 0);
    if (!ok) atomicCAS(p.status, 0, 42);
    for (int it = 0; it < n_local; ++it) {
      const int s = buf_of(it), k = use_of(it);
      ok = ok && ptx::mbar_wait(bar(SB_PF + s), (uint32_t)(k & 1));            // planes[s] written and fenced
      if (k >= 1) ok = ok && ptx::mbar_wait(bar(SB_TE + s), (uint32_t)((k - 1) & 1));   // accumulators s drained
      if (!ok) atomicCAS(p.status, 0, 43);
      ptx::tc_fence_after_sync();
      const uint64_t a_tmpl = ptx::smem_desc(sbase + CF::OFF_PLANES + s * CF::PLANE_BYTES, CF::HALF_K ? 0 : CF::PPAD * 16, CF::PU * 16);
      const uint32_t d_tmem = tmem_base + (uint32_t)(s * CF::ACC_COLS);
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int di = tap / 3, dj = tap % 3;
        // S2: parity sub-plane + position shift; TCONV: source position of the row pixel and output phase
        const int sub_off = CF::S2    ? ((di & 1) * 2 + (dj & 1)) * CF::SUB_BYTES + ((di >> 1) * CF::PU + (dj >> 1)) * 16
                            : CF::S1N ? (di * CF::PU + dj) * 16
                                      : (((di == 2) ? 0 : 1) * CF::PU + ((dj == 2) ? 0 : 1)) * 16;
        const int ph = CF::NPH == 1 ? 0 : ((di == 1) ? 2 : 0) + ((dj == 1) ? 1 : 0);
        const bool first = CF::NPH == 1 ? (tap == 0) : (tap == 0 || tap == 1 || tap == 3 || tap == 4);
#pragma unroll
        for (int kk = 0; kk < CF::KS; ++kk) {
          const uint32_t a_off = (uint32_t)(sub_off + kk * 2 * CF::PPAD * 16) >> 4;
          const uint32_t b_off = (uint32_t)((tap * CF::KS + kk) * CF::N * 32) >> 4;
#pragma unroll
          for (int j = 0; j < J; ++j)
            if (leader) ptx::mma_bf16_ss(d_tmem + (uint32_t)((j * CF::NPH + ph) * CF::N), a_tmpl + a_off + (uint32_t)(8 * j),
                                         b_tmpl + b_off, idesc, (first && kk == 0) ? 0u : 1u);
        }
      }
      if (leader) ptx::mma_commit(bar(SB_TF + s));
      __syncwarp();
    }
  } else if (warp < 4) {
    // =========================== front: stage -> planes ===========================
    for (int it = 0; it < n_local; ++it) {
      const int s = buf_of(it), k = use_of(it);
      ok = ok && ptx::mbar_wait(bar(SB_X + s), (uint32_t)(k & 1));
      if (k >= 1) ok = ok && ptx::mbar_wait(bar(SB_TF + s), (uint32_t)((k - 1) & 1));   // planes[s] no longer read
      if (!ok) atomicCAS(p.status, 0, 44);
      transform_tile<CF>(sbase + CF::OFF_STAGE + s * CF::STAGE_BYTES, sbase + CF::OFF_PLANES + s * CF::PLANE_BYTES, tid);
      ptx::mbar_arrive(bar(SB_XF + s));
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(bar(SB_PF + s));
    }
  } else {
    // =========================== back: 8 epilogue warps ===========================
    const int wb = warp - 4;
    const int q = wb & 3, half = wb >> 2;
    const int r = q * 32 + lane;
    const int v = r >> 3, u = r & 7;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    // work items of a tile: (sub-tile j, phase ph); the two halves take alternate items
    constexpr int ITEMS = J * CF::NPH;
    for (int it = 0; it < n_local; ++it) {
      const int s = buf_of(it), k = use_of(it);
      int b, y0, x0;
      tile_coords((int)blockIdx.x + it * (int)gridDim.x, b, y0, x0);
      const int ry = y0 + v;
      // adjoint epilogue: fetch this thread's z (and old y) vectors while the MMAs still run
      constexpr bool ADJ = CF::S1N || (MODE == MODE_TCONV && R == 2);     // kernels with the prologue-adjoint epilogue
      constexpr int NSLOT = ADJ ? ((ITEMS + 1) / 2) * (CF::NREAL / 16 > 0 ? CF::NREAL / 16 : 1) : 1;
      uint4 pre_z[NSLOT], pre_y[NSLOT];
      if (ADJ && p.adj) {
        constexpr int CY = CF::NREAL / 2;
#pragma unroll
        for (int w0 = 0; w0 < ITEMS; w0 += 2) {
          const int item = w0 + half;
          if (item >= ITEMS) break;
          const int pj = item / CF::NPH, pph = item % CF::NPH;
          const int rx = x0 + 8 * pj + u;
          const int poy = CF::NPH == 1 ? ry : 2 * ry + (pph >> 1);
          const int pox = CF::NPH == 1 ? rx : 2 * rx + (pph & 1);
          const size_t pix = ((size_t)b * p.OH + poy) * p.OW + pox;
#pragma unroll
          for (int cc = 0; cc < CY / 8; ++cc) {
            const int slot = (w0 / 2) * (CY / 8) + cc;
            pre_z[slot] = make_uint4(0u, 0u, 0u, 0u);
            pre_y[slot] = make_uint4(0u, 0u, 0u, 0u);
            if (ry < p.RH && rx < p.RW) {
              pre_z[slot] = __ldg(reinterpret_cast<const uint4*>(p.res + pix * CY + cc * 8));
              if (p.adj_accumulate) pre_y[slot] = *reinterpret_cast<const uint4*>(p.y + pix * CY + cc * 8);
            }
          }
        }
      }
      ok = ok && ptx::mbar_wait(bar(SB_TF + s), (uint32_t)(k & 1));
      if (!ok) atomicCAS(p.status, 0, 45);
      ptx::tc_fence_after_sync();
#pragma unroll
      for (int w0 = 0; w0 < ITEMS; w0 += 2) {
        const int item = w0 + half;
        if (ITEMS == 1 && half == 1) break;
        if (item >= ITEMS) break;
        const int j = item / CF::NPH, ph = item % CF::NPH;
        const int rx = x0 + 8 * j + u;
        const bool live = ok && ry < p.RH && rx < p.RW;
        const int oy = CF::NPH == 1 ? ry : 2 * ry + (ph >> 1);
        const int ox = CF::NPH == 1 ? rx : 2 * rx + (ph & 1);
        const size_t pix = ((size_t)b * p.OH + oy) * p.OW + ox;
        const uint32_t col0 = lane_base + (uint32_t)(s * CF::ACC_COLS + (j * CF::NPH + ph) * CF::N);
        if (ADJ && p.adj) {
          // y[pix][c] (+)= dA0[c] f'(z[c]) - dA1[c] f'(-z[c]) with [dA0 | dA1] = the two accumulator halves and the
          // forward's dropout mask on dA (flow_architecture.py:14-29, 143-145); z / old y were fetched before the wait
          const int kind = p.adj - 1;
          constexpr int CY = CF::NREAL / 2;
#pragma unroll
          for (int cc = 0; cc < CY / 8; ++cc) {
            const int c0 = cc * 8;
            const int slot = (w0 / 2) * (CY / 8) + cc;
            float a0[8], a1[8];
            ptx::tmem_ld8x2(col0 + c0, col0 + CY + c0, a0, a1);
            if (live) {
              bf16x8 xv, old, st;
              xv.v = pre_z[slot];
              old.v = pre_y[slot];
              float sc0[8], sc1[8];
              const bool drop = CF::S1N && p.keep_prob < 1.0f;     // the stride-2 convs have no dropout on their input
              if (drop) {
                const unsigned long long seed = p.seed + (p.seed_dev ? *p.seed_dev : 0ull);
                dropout_scale8(seed, (unsigned long long)pix * CF::NREAL + c0, p.keep_prob, sc0);
                dropout_scale8(seed, (unsigned long long)pix * CF::NREAL + CY + c0, p.keep_prob, sc1);
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float xf = __bfloat162float(xv.h[i]);
                // the two-step path stores dA as bf16 before the adjoint; keep that rounding point
                float f0 = __bfloat162float(__float2bfloat16(a0[i] + sbias[c0 + i]));
                float f1 = __bfloat162float(__float2bfloat16(a1[i] + sbias[CY + c0 + i]));
                if (drop) { f0 *= sc0[i]; f1 *= sc1[i]; }
                float d;
                if (kind == FN_PRO_CONCAT_ELU) {
                  const float e = __expf(-fabsf(xf));
                  d = xf > 0.f ? f0 - f1 * e : f0 * e - f1;
                } else {   // FN_PRO_CONCAT_RELU
                  d = (xf > 0.f ? f0 : 0.f) - (xf < 0.f ? f1 : 0.f);
                }
                if (p.adj_accumulate) d += __bfloat162float(old.h[i]);
                st.h[i] = __float2bfloat16(d);
              }
              *reinterpret_cast<uint4*>(p.y + pix * CY + c0) = st.v;
            }
          }
          continue;
        }
#pragma unroll
        for (int c0 = 0; c0 < CF::NREAL; c0 += 8) {
          float o[8];
          ptx::tmem_ld8(col0 + c0, o);
          if (live) {
            bf16x8 st;
#pragma unroll
            for (int i = 0; i < 8; ++i) st.h[i] = __float2bfloat16(o[i] + sbias[c0 + i]);
            if (p.y) *reinterpret_cast<uint4*>(p.y + pix * CF::NREAL + c0) = st.v;
            if (p.yp) {   // [B][2 NREAL/8][OH][OW][8] = concat_elu of the bf16 output
              uint4 e0, e1;
              concat_elu8(st.v, e0, e1);
              const size_t plane_px = (size_t)p.OH * p.OW;
              __nv_bfloat16* d0 = p.yp + (((size_t)b * (CF::NREAL / 4) + c0 / 8) * plane_px + (size_t)oy * p.OW + ox) * 8;
              *reinterpret_cast<uint4*>(d0) = e0;
              *reinterpret_cast<uint4*>(d0 + (size_t)(CF::NREAL / 8) * plane_px * 8) = e1;
            }
          }
        }
      }
      ptx::tc_fence_before_sync();
      ptx::mbar_arrive(bar(SB_TE + s));
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == WARP_PRODUCER) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, (uint32_t)CF::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------ host side

struct AdjArgs {
  int adj = 0, accumulate = 0;
  const void* res = nullptr;
  float keep_prob = 1.0f;
  unsigned long long seed = 0;
  const unsigned long long* seed_dev = nullptr;
  void* y_planes = nullptr;
};

template <int MODE, int CLOG, int J, int NBUF, int R = 1>
static int launch(const void* x, const void* wpk, const float* bias, void* y, int B, int H, int W, int Cout, int* status,
                  cudaStream_t stream, const AdjArgs& adj = AdjArgs()) {
  using CF = Cfg<MODE, CLOG, J, NBUF, R>;
  Params p;
  memset(&p, 0, sizeof(p));
  if (!make_map_nhwc(&p.tm_x, x, B, H, W, CF::C, CF::SW_, CF::SH_)) { set_error("conv_s2t: tensor map failed"); return FN_E_UNSUPPORTED; }
  p.wpk = (const __nv_bfloat16*)wpk;
  p.bias = bias;
  p.y = (__nv_bfloat16*)y;
  p.yp = (__nv_bfloat16*)adj.y_planes;
  p.status = status;
  p.B = B; p.H = H; p.W = W;
  if (CF::S2) { p.OH = H / 2; p.OW = W / 2; p.RH = p.OH; p.RW = p.OW; }
  else if (CF::S1N) { p.OH = H; p.OW = W; p.RH = H; p.RW = W; }
  else { p.OH = 2 * H; p.OW = 2 * W; p.RH = H; p.RW = W; }
  p.tiles_u = (p.RW + 8 * J - 1) / (8 * J);
  p.tiles_v = (p.RH + 15) / 16;
  p.ntiles = B * p.tiles_u * p.tiles_v;
  p.Cout = Cout;
  p.adj = adj.adj;
  p.adj_accumulate = adj.accumulate;
  p.res = (const __nv_bfloat16*)adj.res;
  p.keep_prob = adj.keep_prob;
  p.seed = adj.seed;
  p.seed_dev = adj.seed_dev;
  auto kern = k_conv_s2t<MODE, CLOG, J, NBUF, R>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(s2t): %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  int per_sm = (227 * 1024) / (CF::SMEM + 1024);
  const int tmem_lim = 512 / CF::TMEM_COLS;
  if (per_sm > tmem_lim) per_sm = tmem_lim;
  if (per_sm > 4) per_sm = 4;
  if (per_sm < 1) per_sm = 1;
  int grid = 148 * per_sm;
  if (grid > p.ntiles) grid = p.ntiles;
  launch_pdl(kern, dim3(grid), dim3(THREADS), (size_t)CF::SMEM, stream, p);
  count_launch();
  return check_launch("k_conv_s2t");
}

}  // namespace s2t

int* tc_status_word();

// stride-2 conv_1 of a down-sampling block
bool conv_s2p_try(const fn_conv_desc& d, cudaStream_t stream, bool dry_run, int& rc) {
  using namespace s2t;
  rc = 0;
  if (d.ksize != 3 || d.stride != 2 || d.prologue != FN_PRO_CONCAT_ELU || d.epilogue != FN_EPI_BIAS || d.skip ||
      d.keep_prob < 1.0f || d.w_tc == nullptr)
    return false;
  if ((d.H & 1) || (d.W & 1) || d.H < 32) return false;       // TF pads (0,1) on even sizes only; short images: generic kernel
  if (d.Cout != 2 * d.Cin) return false;
  if (get_encode() == nullptr) return false;
  const int clog = d.Cin == 8 ? 0 : d.Cin == 16 ? 1 : d.Cin == 32 ? 2 : -1;
  if (clog < 0) return false;
  const int OW = d.W / 2;
#define S2_CASE(CL, JJ, NB, CAP)                                                                              \
  if (clog == CL && ((OW % (8 * JJ)) == 0 || JJ == 1) && Cfg<MODE_S2, CL, JJ, NB>::SMEM <= CAP * 1024) {       \
    if (!dry_run) {                                                                                           \
      int* st = tc_status_word();                                                                             \
      AdjArgs out;                                                                                            \
      out.y_planes = d.y_planes;                                                                              \
      rc = st ? launch<MODE_S2, CL, JJ, NB>(d.x, d.w_tc, d.bias, d.y, d.B, d.H, d.W, d.Cout, st, stream, out) \
              : (int)cudaErrorMemoryAllocation;                                                               \
    }                                                                                                         \
    return true;                                                                                              \
  }
  S2_CASE(0, 2, 2, 112) S2_CASE(0, 1, 2, 112)
  S2_CASE(1, 2, 2, 112) S2_CASE(1, 1, 2, 150)
  S2_CASE(2, 1, 1, 200)
#undef S2_CASE
  return false;
}

// stride-2 3x3 conv without prologue, Cout = 2 Cin: the data gradient of a transposed conv (conv2d_transpose is defined as
// the gradient of this conv, flow_architecture.py:84-98), raw dY in, K = Cin
bool conv_s2n_try(const fn_conv_desc& d, cudaStream_t stream, bool dry_run, int& rc) {
  using namespace s2t;
  rc = 0;
  if (d.ksize != 3 || d.stride != 2 || d.prologue != FN_PRO_NONE || d.epilogue != FN_EPI_BIAS || d.skip || d.w_tc == nullptr)
    return false;
  if ((d.H & 1) || (d.W & 1) || d.H < 32) return false;
  if (d.Cout != 2 * d.Cin) return false;
  if (get_encode() == nullptr) return false;
  const int clog = d.Cin == 8 ? 0 : d.Cin == 16 ? 1 : d.Cin == 32 ? 2 : -1;
  if (clog < 0) return false;
  const int OW = d.W / 2;
#define S2N_CASE(CL, JJ, NB, CAP)                                                                             \
  if (clog == CL && ((OW % (8 * JJ)) == 0 || JJ == 1) && Cfg<MODE_S2N, CL, JJ, NB>::SMEM <= CAP * 1024) {      \
    if (!dry_run) {                                                                                           \
      int* st = tc_status_word();                                                                             \
      rc = st ? launch<MODE_S2N, CL, JJ, NB>(d.x, d.w_tc, d.bias, d.y, d.B, d.H, d.W, d.Cout, st, stream)     \
              : (int)cudaErrorMemoryAllocation;                                                               \
    }                                                                                                         \
    return true;                                                                                              \
  }
  S2N_CASE(0, 2, 2, 112) S2N_CASE(0, 1, 2, 112)
  S2N_CASE(1, 2, 2, 112) S2N_CASE(1, 1, 2, 112)
  S2N_CASE(2, 2, 2, 150) S2N_CASE(2, 1, 2, 150) S2N_CASE(2, 1, 1, 200)
#undef S2N_CASE
  return false;
}

// stride-1 3x3 conv without prologue, Cout = C or 2C: the data-gradient convs of the backward pass (flipped weights)
bool conv_s1n_try(const fn_conv_desc& d, cudaStream_t stream, bool dry_run, int& rc) {
  using namespace s2t;
  rc = 0;
  const bool adjoint = d.epilogue == FN_EPI_PRO_BWD;
  if (d.ksize != 3 || d.stride != 1 || d.prologue != FN_PRO_NONE || (d.epilogue != FN_EPI_BIAS && !adjoint) || d.skip ||
      d.w_tc == nullptr || d.H < 16)
    return false;
  AdjArgs adj;
  if (adjoint) {
    if (d.adj_prologue != FN_PRO_CONCAT_ELU && d.adj_prologue != FN_PRO_CONCAT_RELU) return false;   // others: two launches
    const int cy = d.Cout / 2;
    if (cy % 8) return false;
    if (d.res == nullptr || d.res_pool || d.Cres != cy || d.RH != d.H || d.RW != d.W || d.dropout_offset != 0) return false;
    adj.adj = 1 + d.adj_prologue;
    adj.accumulate = d.adj_accumulate;
    adj.res = d.res;
    adj.keep_prob = d.keep_prob;
    adj.seed = d.dropout_seed;
    adj.seed_dev = (const unsigned long long*)d.dropout_seed_dev;
  }
  if (get_encode() == nullptr) return false;
  const int clog = d.Cin == 8 ? 0 : d.Cin == 16 ? 1 : d.Cin == 32 ? 2 : d.Cin == 64 ? 3 : -1;
  const int r = d.Cout == d.Cin ? 1 : d.Cout == 2 * d.Cin ? 2 : 0;
  if (clog < 0 || r == 0) return false;
#define S1N_CASE(CL, RR, JJ, NB, CAP)                                                                            \
  if (clog == CL && r == RR && ((d.W % (8 * JJ)) == 0 || JJ == 1) && Cfg<MODE_S1N, CL, JJ, NB, RR>::SMEM <= CAP * 1024) { \
    if (!dry_run) {                                                                                              \
      int* st = tc_status_word();                                                                                \
      rc = st ? launch<MODE_S1N, CL, JJ, NB, RR>(d.x, d.w_tc, d.bias, d.y, d.B, d.H, d.W, d.Cout, st, stream, adj) \
              : (int)cudaErrorMemoryAllocation;                                                                  \
    }                                                                                                            \
    return true;                                                                                                 \
  }
  S1N_CASE(0, 1, 4, 2, 112) S1N_CASE(0, 1, 2, 2, 112) S1N_CASE(0, 1, 1, 2, 112)
  S1N_CASE(0, 2, 4, 2, 112) S1N_CASE(0, 2, 2, 2, 112) S1N_CASE(0, 2, 1, 2, 112)
  S1N_CASE(1, 1, 4, 2, 112) S1N_CASE(1, 1, 2, 2, 112) S1N_CASE(1, 1, 1, 2, 112)
  S1N_CASE(1, 2, 4, 2, 112) S1N_CASE(1, 2, 2, 2, 112) S1N_CASE(1, 2, 1, 2, 112)
  S1N_CASE(2, 1, 2, 2, 112) S1N_CASE(2, 1, 1, 2, 112)
  S1N_CASE(2, 2, 2, 2, 150) S1N_CASE(2, 2, 1, 2, 150)
  S1N_CASE(3, 1, 2, 2, 200) S1N_CASE(3, 1, 1, 2, 200)
#undef S1N_CASE
  return false;
}

bool tconv_s2p_try(const fn_tconv_desc& d, cudaStream_t stream, bool dry_run, int& rc) {
  using namespace s2t;
  rc = 0;
  const int r = d.Cout * 2 == d.Cin ? 1 : d.Cout == d.Cin ? 2 : 0;
  if (d.w_tc == nullptr || r == 0 || d.H < 16) return false;
  AdjArgs adj;
  if (d.epilogue == FN_EPI_PRO_BWD) {
    if (r != 2 || d.res == nullptr || (d.adj_prologue != FN_PRO_CONCAT_ELU && d.adj_prologue != FN_PRO_CONCAT_RELU)) return false;
    adj.adj = 1 + d.adj_prologue;
    adj.accumulate = d.adj_accumulate;
    adj.res = d.res;
  }
  adj.y_planes = d.y_planes;
  if (adj.adj && d.y_planes) return false;
  if (get_encode() == nullptr) return false;
  const int clog = d.Cin == 16 ? 1 : d.Cin == 32 ? 2 : d.Cin == 64 ? 3 : -1;
  if (clog < 0) return false;
#define TC_CASE(CL, RR, JJ, NB, CAP)                                                                          \
  if (clog == CL && r == RR && ((d.W % (8 * JJ)) == 0 || JJ == 1) && !(adj.adj && JJ > 1 && CL == 1) &&        \
      Cfg<MODE_TCONV, CL, JJ, NB, RR>::SMEM <= CAP * 1024) {                                                  \
    if (!dry_run) {                                                                                           \
      int* st = tc_status_word();                                                                             \
      rc = st ? launch<MODE_TCONV, CL, JJ, NB, RR>(d.x, d.w_tc, d.bias, d.y, d.B, d.H, d.W, d.Cout, st, stream, adj) \
              : (int)cudaErrorMemoryAllocation;                                                               \
    }                                                                                                         \
    return true;                                                                                              \
  }
  TC_CASE(1, 1, 4, 2, 112) TC_CASE(1, 1, 2, 2, 112) TC_CASE(1, 1, 1, 2, 112)
  TC_CASE(2, 1, 2, 2, 112) TC_CASE(2, 1, 1, 2, 112)
  TC_CASE(3, 1, 2, 2, 112) TC_CASE(3, 1, 1, 2, 150)
  // Cout = Cin: data gradients of the stride-2 convs (dY of 2C channels -> d concat_elu(x) of 2C channels)
  TC_CASE(1, 2, 2, 2, 112) TC_CASE(1, 2, 1, 2, 112)
  TC_CASE(2, 2, 2, 2, 150) TC_CASE(2, 2, 1, 2, 150)
  TC_CASE(3, 2, 1, 2, 200)
#undef TC_CASE
  return false;
}

}  // namespace fn
