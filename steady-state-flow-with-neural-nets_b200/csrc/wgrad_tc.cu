// Weight gradient of the 'SAME' convolutions on the tcgen05 tensor cores (what autodiff derives for
// tf.nn.conv2d / conv2d_transpose / the nin matmul, reference flow_architecture.py:64,78,101 under
// flow_net.train, model/flow_net.py:44-46):
//
//     dW[tap][k][n] = sum_pix  pro(x)[s*pix + tap - pad][k] * dY[pix][n]
//
// is, per tap, a GEMM whose CONTRACTION runs over pixels.  Both operands live in shared memory exactly
// as the forward kernels keep them -- 8-channel planes `[chunk][pixel][8]`, 16 B per pixel -- which is
// the canonical MN-major (transposed) no-swizzle UMMA layout: 8 consecutive pixels x 8 channels form
// one 128-byte core matrix, the next 8 pixels follow at LBO = 128 B and the next channel chunk at
// SBO = plane stride.  One MMA (M = 128 activation channels, N = Cout, K = 16) therefore consumes 16
// consecutive pixels of one tile row, and a filter tap is once more only a start-address offset into
// the haloed activation planes (stride 2: parity sub-planes, as in conv_s2t.cu).
//
// Narrow stride-1 3x3 layers (Keff <= 32) fold the nine taps into ONE instruction per 16 pixels: the three filter
// columns as three column-shifted copies of the activation planes stacked along M (a3), the three filter rows as the
// dY rows r-2, r-1, r stacked along N -- the dY tile is loaded as [row][chunk][x][8], which makes the chunks of
// consecutive rows equidistant, so one B descriptor spans them (n3).
// Stride-2 layers with Keff <= 32 start every instruction in parity sub-plane 0 and let its M rows run through all
// four sub-planes (s4): four accumulators (tap pair offsets) instead of nine.
//
// Persistent CTAs: CTA (g, s) owns one group g of taps / one 128-channel block of K and every S-th
// pixel tile; its accumulators stay in TMEM for the whole kernel and are written once, as an fp32
// partial `[s][tap][k][n]`; a second kernel adds the S partials in a fixed order (deterministic).
// warps 0-3: prologue transform (stage -> planes), later the TMEM read-out; warp 4: TMA producer and
// TMEM owner; warp 5: MMA issuer.
#include "common.cuh"
#include "ptx.cuh"
#include "tma.cuh"

namespace fn {
using namespace tma;
namespace wg {

constexpr int TV = 8;                          // pixel tile of dY: 8 rows of TU = 16 (or 32) pixels = 8 (16) K-steps
constexpr int THREADS = 192;
constexpr int WARP_PRODUCER = 4, WARP_ISSUER = 5;
constexpr int HDR = 1024;           // mbarriers [0,256), TMEM slot at 256, 256 B of bf16 ones at 512
constexpr int OFF_ONES = 512;
constexpr int MAXBUF = 8;
enum { MODE_S1 = 0, MODE_S2 = 1 };

template <int MODE, int TU = 16>
struct Geo {
  static constexpr bool S2 = MODE == MODE_S2;
  static constexpr int TUX = TU;
  static constexpr int KSU = TU / 16;                                                                     // K-steps per tile row
  static constexpr int DY_PLANE = TV * TU * 16;                                                           // bytes of one 8-channel dY plane
  static constexpr int SW_ = S2 ? 2 * TU + 1 : TU + 2, SH_ = S2 ? 2 * TV + 1 : TV + 2, SP = SW_ * SH_;   // stage tile (TMA box)
  static constexpr int PU = S2 ? TU + 1 : TU + 2, PV = S2 ? TV + 1 : TV + 2, P = PU * PV;                // one (sub-)plane
  static constexpr int PPAD = (P & 1) ? P : P + 1;                                                       // odd: planes spread over banks
  static constexpr int NSUB = S2 ? 4 : 1;
  static constexpr int ROW_BYTES = PU * 16;                                                              // one K-step down
};

struct Params {
  CUtensorMap tm_x;     // 4-D (C, W, H, B)
  CUtensorMap tm_dy;    // 5-D (8, N/8, OW, OH, B)
  float* partial;       // [S][T][Keff][N]
  float* bias_partial;  // [S][N] or null: db[n] = sum_pix dY[pix][n] as one more accumulator fed by a plane of ones
  int* status;
  int B, H, W, C, OH, OW, N;
  int CH, chshift;      // raw 8-channel chunks of x
  int pro, Keff, T;
  int mblocks;          // blocks of 8*kpl activation channels
  int kpl;              // activation planes one CTA keeps (16, or 8 when a stride-2 tile would not fit)
  int tap_groups, S;
  int tiles_u, tiles_v, ntiles;
  int Npad, nchunks;
  uint32_t idesc;
  int nbuf;
  uint32_t off_a, a_bytes, off_dy, dy_bytes, off_stage, stage_bytes, sub_bytes;
  uint32_t tmem_cols;
  int ksize, swap, m64;
  int nt;               // accumulators per CTA
  int a3;               // 1: three column-shifted copies of every plane, the 3 taps of a filter row share one MMA
  int s4;               // stride 2, Keff <= 32: the four parity sub-planes lie one behind the other with the same chunk stride, so an
                        // M = 64 / 128 instruction started in sub-plane 0 covers all four -- the taps (2a+p, 2b+q) for the four
                        // parities (p, q) at once; four accumulators (a, b) instead of nine, rows = (parity, channel)
  int n3;               // a3 and: the three filter ROWS share the MMA too -- the dY tile lands as [row][chunk][x][8] (one tensor copy), so
                        // the chunks of consecutive rows are equidistant and one B descriptor (chunk stride = one x-extent) reads the
                        // N = 3 * nchunks * 8 columns of dY rows r-2, r-1, r against activation row r: nine taps per instruction,
                        // TV + 2 instructions per K-step column instead of 3 TV (the first / last two activation rows pair with 1 / 2 rows)
  int n3_cols;          // accumulator columns of that MMA; the bias accumulator follows
  uint32_t idesc_rows[3];   // instruction descriptors for 1 / 2 / 3 dY rows
  uint32_t idesc_bias;
  float keep_prob;
  unsigned long long seed;
  const unsigned long long* seed_dev;
};

enum { SB_X = 0, SB_PF = MAXBUF, SB_TF = 2 * MAXBUF, SB_FIN = 3 * MAXBUF };

// instruction descriptor: BF16 x BF16 -> FP32, A and B MN-major (bits 15 / 16)
static uint32_t idesc_mn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// stage [pixel][C] raw bf16 -> this CTA's operand planes: prologue (CE: the single-exponential concat_elu; else the
// generic switch), optional dropout mask of the forward pass, parity split (stride 2), three shifted copies (a3)
template <typename G, bool CE, bool DROP>
__device__ __forceinline__ void transform_tile(const Params& p, uint32_t stage, uint32_t planes, int tid, int items, int mb, int b,
                                               int y0, int x0) {
  const int kpl = p.kpl;
  const int mult = CE ? 2 : pro_mult(p.pro);
#pragma unroll 2
  for (int idx = tid; idx < items; idx += 128) {
    const int ch = idx & (p.CH - 1);
    const int sp = idx >> p.chshift;
    bf16x8 in, o0, o1;
    in.v = lds128(stage + (uint32_t)idx * 16u);
    const int r = sp / G::SW_, c = sp - r * G::SW_;
    float sc0[8], sc1[8];
    if (DROP) {     // tf.nn.dropout on the conv input (flow_architecture.py:144-145): same mask as the forward
      const int iy = G::S2 ? 2 * y0 + r : y0 - 1 + r, ix = G::S2 ? 2 * x0 + c : x0 - 1 + c;
      const unsigned long long pidx = ((unsigned long long)b * p.H + iy) * p.W + ix;
      const unsigned long long seed = p.seed + (p.seed_dev ? *p.seed_dev : 0ull);
      dropout_scale8(seed, pidx * p.Keff + ch * 8, p.keep_prob, sc0);
      if (mult == 2) dropout_scale8(seed, pidx * p.Keff + p.C + ch * 8, p.keep_prob, sc1);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v = __bfloat162float(in.h[j]);
      float a0, a1;
      if (CE) {
        const float t = __expf(-fabsf(v)) - 1.f;
        a0 = v > 0.f ? v : t;
        a1 = v > 0.f ? t : -v;
      } else {
        pro_apply(p.pro, v, a0, a1);
      }
      if (DROP) {     // the forward rounds the prologue output to bf16 before scaling
        a0 = __bfloat162float(__float2bfloat16(a0)) * sc0[j];
        a1 = __bfloat162float(__float2bfloat16(a1)) * sc1[j];
      }
      o0.h[j] = __float2bfloat16(a0);
      o1.h[j] = __float2bfloat16(a1);
    }
    uint32_t dst = planes;
    int pp = sp;
    if (G::S2) {
      dst += (uint32_t)((r & 1) * 2 + (c & 1)) * p.sub_bytes;
      pp = (r >> 1) * G::PU + (c >> 1);
    }
    const int l0 = ch - kpl * mb, l1 = p.CH + ch - kpl * mb;
    if (p.a3) {
      // copy du holds the plane shifted left by du pixels: copy[du][pp] = plane[pp + du]
#pragma unroll
      for (int du = 0; du < 3; ++du) {
        if (pp >= du) {
          sts128(dst + (uint32_t)((du * kpl + l0) * G::PPAD + pp - du) * 16u, o0.v);
          if (mult == 2) sts128(dst + (uint32_t)((du * kpl + l1) * G::PPAD + pp - du) * 16u, o1.v);
        }
      }
    } else {
      if (l0 >= 0 && l0 < kpl) sts128(dst + (uint32_t)(l0 * G::PPAD + pp) * 16u, o0.v);
      if (mult == 2 && l1 >= 0 && l1 < kpl) sts128(dst + (uint32_t)(l1 * G::PPAD + pp) * 16u, o1.v);
    }
  }
}

// N3: the nine-taps-per-instruction form, a kernel of its own so that the other shapes' code is what it was (folded into one
// kernel behind a run-time flag the M = 64 launches of the other shapes ran 15-50 % slower: the issue loop is what bounds them)
template <int MODE, int TU = 16, bool N3 = false>
__global__ void __launch_bounds__(THREADS) k_wgrad_tc(const __grid_constant__ Params p) {
  using G = Geo<MODE, TU>;
  constexpr int DY_PLANE = G::DY_PLANE;
  // no static shared memory in these kernels: the dynamic window starts on a 1024-byte boundary, so the base is a
  // compile-time constant of the shared state space (TMA destinations want 128-byte alignment)
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = ptx::smem_u32(smem);
  auto bar = [&](int i) { return sbase + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + 256);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int tid = threadIdx.x;

  // CTA -> (pixel split s, tap group, K block)
  const int s_idx = (int)blockIdx.x % p.S;
  const int grp = (int)blockIdx.x / p.S;
  const int mb = grp % p.mblocks;
  const int kpl = p.kpl;
  const int NT = p.nt;                              // accumulators (taps, or filter rows with a3) of this CTA
  const int tap0 = (grp / p.mblocks) * NT;
  const int ntaps = min(NT, p.T - tap0);            // the last tap group may be short
  const int n_local = s_idx < p.ntiles ? (p.ntiles - 1 - s_idx) / p.S + 1 : 0;
  auto tile_coords = [&](int t, int& b, int& y0, int& x0) {
    const int tu = t % p.tiles_u;
    const int r = t / p.tiles_u;
    const int tv = r % p.tiles_v;
    b = r / p.tiles_v;
    x0 = tu * TU;
    y0 = tv * TV;
  };
  const int nbuf = p.nbuf;
  auto buf_of = [&](int it) { return it % nbuf; };
  auto use_of = [&](int it) { return it / nbuf; };

  if (warp == WARP_PRODUCER) {
    ptx::tmem_alloc(sbase + 256, p.tmem_cols);
    ptx::tmem_relinquish();
    if (lane == 0) {
      for (int s = 0; s < MAXBUF; ++s) {
        ptx::mbar_init(bar(SB_X + s), 1);
        ptx::mbar_init(bar(SB_PF + s), 128);
        ptx::mbar_init(bar(SB_TF + s), 1);
      }
      ptx::mbar_init(bar(SB_FIN), 1);
      ptx::fence_mbar_init();
    }
    __syncwarp();
  }
  const bool do_bias = p.bias_partial != nullptr && grp == 0;
  const uint32_t n3_rp = (uint32_t)(p.nchunks * TU * 16);        // n3: bytes of one dY row (all chunks)
  if (tid < 16) sts128(sbase + OFF_ONES + (uint32_t)tid * 16u, make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u));
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  bool ok = true;

  if (warp == WARP_PRODUCER) {
    // =========================== producer ===========================
    if (lane == 0) {
      for (int it = 0; it < n_local && ok; ++it) {
        const int s = buf_of(it), k = use_of(it);
        int b, y0, x0;
        tile_coords(s_idx + it * p.S, b, y0, x0);
        if (k >= 1) {
          ok = ptx::mbar_wait(bar(SB_TF + s), (uint32_t)((k - 1) & 1));     // MMAs of the previous use have read planes / dY
          if (!ok) { atomicCAS(p.status, 0, 61); break; }
        }
        const bool skip_x = p.swap & 8, skip_dy = p.swap & 4;     // timing experiments only (results are then garbage)
        ptx::mbar_arrive_expect_tx(bar(SB_X + s), (uint32_t)((skip_x ? 0 : G::SP * p.C * 2) + (skip_dy ? 0 : p.nchunks * DY_PLANE)));
        const uint32_t st = sbase + p.off_stage + (uint32_t)s * p.stage_bytes;
        if (!skip_x) {
          if (G::S2) tma_load_4d(st, &p.tm_x, bar(SB_X + s), 0, 2 * x0, 2 * y0, b);
          else tma_load_4d(st, &p.tm_x, bar(SB_X + s), 0, x0 - 1, y0 - 1, b);
        }
        const uint32_t dy = sbase + p.off_dy + (uint32_t)s * p.dy_bytes;
        if (!skip_dy) {
          if (N3) tma_load_5d(dy, &p.tm_dy, bar(SB_X + s), 0, x0, 0, y0, b);     // (8, x, chunk, y, b): one request per tile
          else
            for (int c = 0; c < p.nchunks; ++c) tma_load_5d(dy + (uint32_t)(c * DY_PLANE), &p.tm_dy, bar(SB_X + s), 0, c, x0, y0, b);
        }
      }
    }
    __syncwarp();
  } else if (warp == WARP_ISSUER) {
    // =========================== issuer ===========================
    const uint32_t leader = ptx::elect_one();
    // byte offsets of this CTA's taps inside the activation planes
    uint32_t tapoff[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int tap = tap0 + t;
      int di = tap / 3, dj = tap % 3;
      if (p.ksize == 1) { di = 1; dj = 1; }
      tapoff[t] = G::S2 ? (uint32_t)((di & 1) * 2 + (dj & 1)) * p.sub_bytes + (uint32_t)(((di >> 1) * G::PU + (dj >> 1)) * 16)
                        : (uint32_t)((di * G::PU + dj) * 16);
      if (p.a3) tapoff[t] = (uint32_t)(t * G::ROW_BYTES);       // accumulator t = filter row t; columns are M groups
      if (p.s4) tapoff[t] = (uint32_t)((((t >> 1) & 1) * G::PU + (t & 1)) * 16);      // accumulator t = (a, b): offset inside every sub-plane
    }
    const bool swp = p.swap & 1;
    const uint32_t lbo_a = swp ? (uint32_t)(G::PPAD * 16) : 128u, sbo_a = swp ? 128u : (uint32_t)(G::PPAD * 16);
    const uint32_t lbo_b = swp ? (uint32_t)DY_PLANE : 128u, sbo_b = swp ? 128u : (uint32_t)DY_PLANE;
    const uint32_t do_mma = (p.swap & 2) ? 0u : leader;
    for (int it = 0; it < n_local; ++it) {
      const int s = buf_of(it), k = use_of(it);
      ok = ok && ptx::mbar_wait(bar(SB_X + s), (uint32_t)(k & 1));
      ok = ok && ptx::mbar_wait(bar(SB_PF + s), (uint32_t)(k & 1));
      if (!ok) atomicCAS(p.status, 0, 62);
      ptx::tc_fence_after_sync();
      const uint64_t a_tmpl = ptx::smem_desc(sbase + p.off_a + (uint32_t)s * p.a_bytes, lbo_a, sbo_a);
      const uint64_t b_tmpl = ptx::smem_desc(sbase + p.off_dy + (uint32_t)s * p.dy_bytes, lbo_b, sbo_b);
      if (N3) {
        // activation row r (image row y0 - 1 + r) pairs with the dY tile rows r-2 .. r (filter rows 2 .. 0) that exist; column
        // block j of the accumulator = filter row 2 - j.  The full rows come first so that the very first instruction
        // initialises all columns.
        const uint64_t b3 = ptx::smem_desc(sbase + p.off_dy + (uint32_t)s * p.dy_bytes, swp ? (uint32_t)(TU * 16) : 128u,
                                           swp ? 128u : (uint32_t)(TU * 16));
        const uint32_t blk = (uint32_t)(p.nchunks * 8);
#pragma unroll
        for (int rr = 0; rr < TV + 2; ++rr) {
          const int r = rr < TV - 2 ? rr + 2 : (rr < TV ? rr - (TV - 2) : rr);       // 2 .. TV-1, 0, 1, TV, TV+1
          const int lo = r < 2 ? 2 - r : 0, hi = r > TV - 1 ? TV + 1 - r : 2;        // block range [lo, hi]
          const int t0 = r + lo - 2;                                                 // first dY tile row
#pragma unroll
          for (int ub = 0; ub < G::KSU; ++ub)
            if (do_mma)
              ptx::mma_bf16_ss(tmem_base + (uint32_t)lo * blk, a_tmpl + (uint32_t)((r * G::ROW_BYTES + ub * 256) >> 4),
                               b3 + (uint32_t)(((uint32_t)t0 * n3_rp + (uint32_t)ub * 256u) >> 4), p.idesc_rows[hi - lo],
                               (it == 0 && rr == 0 && ub == 0) ? 0u : 1u);
        }
        if (do_bias) {
          const uint64_t ones = ptx::smem_desc(sbase + OFF_ONES, 128u, 0u);
#pragma unroll
          for (int v = 0; v < TV; ++v)
#pragma unroll
            for (int ub = 0; ub < G::KSU; ++ub)
              if (do_mma)
                ptx::mma_bf16_ss(tmem_base + (uint32_t)p.n3_cols, ones, b3 + (uint32_t)(((uint32_t)v * n3_rp + (uint32_t)ub * 256u) >> 4),
                                 p.idesc_bias, (it == 0 && v == 0 && ub == 0) ? 0u : 1u);
        }
        if (leader) ptx::mma_commit(bar(SB_TF + s));
        __syncwarp();
        continue;
      }
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        if (t >= ntaps) break;
        const uint64_t a_tap = a_tmpl + (uint64_t)(tapoff[t] >> 4);
        const uint32_t d_col = tmem_base + (uint32_t)(t * p.Npad);
#pragma unroll
        for (int v = 0; v < TV; ++v)
#pragma unroll
          for (int ub = 0; ub < G::KSU; ++ub)
            if (do_mma)
              ptx::mma_bf16_ss(d_col, a_tap + (uint32_t)((v * G::ROW_BYTES + ub * 256) >> 4),
                               b_tmpl + (uint32_t)(((v * TU + ub * 16) * 16) >> 4), p.idesc, (it == 0 && v == 0 && ub == 0) ? 0u : 1u);
      }
      if (do_bias) {
        // bias gradient: A = 16 pixels x 8 "channels" of ones (every M group reads the same 256 bytes, SBO = 0), so every
        // accumulator row is sum_pix dY[pix][:]; dY is zero outside the image (TMA fill), no masking needed
        const uint64_t ones = ptx::smem_desc(sbase + OFF_ONES, 128u, 0u);
        const uint32_t d_col = tmem_base + (uint32_t)(NT * p.Npad);
#pragma unroll
        for (int v = 0; v < TV; ++v)
#pragma unroll
          for (int ub = 0; ub < G::KSU; ++ub)
            if (do_mma)
              ptx::mma_bf16_ss(d_col, ones, b_tmpl + (uint32_t)(((v * TU + ub * 16) * 16) >> 4), p.idesc,
                               (it == 0 && v == 0 && ub == 0) ? 0u : 1u);
      }
      if (leader) ptx::mma_commit(bar(SB_TF + s));
      __syncwarp();
    }
    if (n_local > 0 && leader) ptx::mma_commit(bar(SB_FIN));
    __syncwarp();
  } else {
    // =========================== front: stage -> planes, then read-out ===========================
    const int items = G::SP * p.CH;
    const int mult = pro_mult(p.pro);
    for (int it = 0; it < n_local; ++it) {
      const int s = buf_of(it), k = use_of(it);
      int b, y0, x0;
      tile_coords(s_idx + it * p.S, b, y0, x0);
      ok = ok && ptx::mbar_wait(bar(SB_X + s), (uint32_t)(k & 1));
      if (!ok) atomicCAS(p.status, 0, 63);
      const uint32_t stage = sbase + p.off_stage + (uint32_t)s * p.stage_bytes;
      const uint32_t planes = sbase + p.off_a + (uint32_t)s * p.a_bytes;
      if (!(p.swap & 16)) {
        const bool ce = p.pro == FN_PRO_CONCAT_ELU;
        if (p.keep_prob < 1.0f) {
          if (ce) transform_tile<G, true, true>(p, stage, planes, tid, items, mb, b, y0, x0);
          else transform_tile<G, false, true>(p, stage, planes, tid, items, mb, b, y0, x0);
        } else {
          if (ce) transform_tile<G, true, false>(p, stage, planes, tid, items, mb, b, y0, x0);
          else transform_tile<G, false, false>(p, stage, planes, tid, items, mb, b, y0, x0);
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(bar(SB_PF + s));
    }
    // ---- read-out: TMEM lane = activation channel (row k), column = tap-local Cout index
    if (n_local > 0) {
      ok = ok && ptx::mbar_wait(bar(SB_FIN), 0);
      if (!ok) atomicCAS(p.status, 0, 64);
      ptx::tc_fence_after_sync();
    }
    // M = 128: accumulator row r lives in TMEM lane r; M = 64 (kpl <= 8): row r in lane 32*(r/16) + r%16 (measured,
    // tools/m64_layout.py)
    const int row_local = p.m64 ? (lane < 16 ? warp * 16 + lane : 1 << 20) : warp * 32 + lane;
    int krow = mb * kpl * 8 + row_local;
    int du_row = 0;                                   // a3: accumulator row = (column shift du, channel k)
    if (p.a3 || p.s4) {                                // s4: du_row = parity sub-plane
      du_row = row_local / (kpl * 8);
      krow = row_local - du_row * kpl * 8;
    }
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int bias_col = N3 ? p.n3_cols : NT * p.Npad;
    if (do_bias && warp == 0) {          // accumulator row 0 = TMEM lane 0 for M = 128 and M = 64 alike
      for (int n0 = 0; n0 < p.N; n0 += 8) {
        float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (n_local > 0) ptx::tmem_ld8(tmem_base + (uint32_t)(bias_col + n0), o);
        if (lane == 0 && ok) {
          float* dst = p.bias_partial + (size_t)s_idx * p.N + n0;
          *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
        }
      }
    }
    if (N3) {
      // column block j (dY rows r-2+j) holds filter row 2 - j
      for (int j = 0; j < 3; ++j) {
        const int tap = 3 * (2 - j) + du_row;
        float* dst = p.partial + (((size_t)s_idx * p.T + tap) * p.Keff + krow) * p.N;
        for (int n0 = 0; n0 < p.N; n0 += 8) {
          float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (n_local > 0) ptx::tmem_ld8(lane_base + (uint32_t)(j * p.nchunks * 8 + n0), o);
          if (du_row < 3 && krow < p.Keff && ok) {
            *reinterpret_cast<float4*>(dst + n0) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4*>(dst + n0 + 4) = make_float4(o[4], o[5], o[6], o[7]);
          }
        }
      }
    } else {
#pragma unroll 1
    for (int t = 0; t < ntaps; ++t) {
      int tap = p.a3 ? 3 * t + du_row : tap0 + t;
      bool row_ok = p.a3 ? du_row < 3 : row_local < kpl * 8;
      if (p.s4) {
        const int di = 2 * (t >> 1) + (du_row >> 1), dj = 2 * (t & 1) + (du_row & 1);
        row_ok = du_row < 4 && di < 3 && dj < 3;
        tap = row_ok ? di * 3 + dj : 0;
      }
      float* dst = p.partial + (((size_t)s_idx * p.T + tap) * p.Keff + krow) * p.N;
      for (int n0 = 0; n0 < p.N; n0 += 8) {
        float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (n_local > 0) ptx::tmem_ld8(lane_base + (uint32_t)(t * p.Npad + n0), o);
        if (row_ok && krow < p.Keff && ok) {
          *reinterpret_cast<float4*>(dst + n0) = make_float4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<float4*>(dst + n0 + 4) = make_float4(o[4], o[5], o[6], o[7]);
        }
      }
    }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == WARP_PRODUCER) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------ host side

static int g_swap = 0;

struct Plan {
  int mode, NT, S, grid, smem, tu;
  bool bias_fits;
  Params p;
};

// x: bf16 [B,H,W,C] (prologue applied on the fly), dY: bf16 [B,OH,OW,N]; result partial [S][T][Keff][N]
static bool plan(Plan& pl, int B, int H, int W, int C, int OH, int OW, int N, int ksize, int stride, int pro) {
  Params& p = pl.p;
  memset(&p, 0, sizeof(p));
  if (get_encode() == nullptr) return false;
  if ((C & 7) || (N & 7) || C > 128 || N > 256 || (C & (C - 1))) return false;
  if (ksize != 1 && ksize != 3) return false;
  if (stride == 2 && (ksize != 3 || (H & 1) || (W & 1))) return false;
  if (stride != 1 && stride != 2) return false;
  pl.mode = stride == 2 ? MODE_S2 : MODE_S1;
  p.B = B; p.H = H; p.W = W; p.C = C; p.OH = OH; p.OW = OW; p.N = N;
  p.CH = C / 8;
  p.chshift = 0;
  while ((1 << p.chshift) < p.CH) ++p.chshift;
  p.pro = pro;
  p.Keff = C * pro_mult(pro);
  p.T = ksize * ksize;
  p.ksize = ksize;
  p.Npad = (N + 15) / 16 * 16;
  p.nchunks = N / 8;
  p.idesc = idesc_mn(128, p.Npad);
  // narrow stride-1 3x3 layers: stack the three column taps along M (3 shifted plane copies), 3 MMAs per K-step
  // instead of 9 -- the MMA rate is set by the 64/128 operand rows read from shared memory, mostly padding here
  p.a3 = (pl.mode == MODE_S1 && ksize == 3 && p.Keff <= 32 && !(g_swap & 64)) ? 1 : 0;
  // taps per CTA: as many accumulators as TMEM holds next to the bias accumulator (every tap group re-loads and
  // re-transforms the tile, so fewer groups is what counts), evened out over the groups; N = 256 gives the bias up
  p.s4 = (pl.mode == MODE_S2 && ksize == 3 && p.Keff <= 32 && !(g_swap & 2048)) ? 1 : 0;
  if (p.a3) {
    pl.NT = 3;
    p.tap_groups = 1;
  } else if (p.s4) {
    pl.NT = 4;
    p.tap_groups = 1;
  } else {
    int cap = 512 / p.Npad - 1;
    if (cap < 2 && p.T > 1) cap = 512 / p.Npad;
    if (cap > p.T) cap = p.T;
    if (cap < 1) cap = 1;
    p.tap_groups = (p.T + cap - 1) / cap;
    pl.NT = (p.T + p.tap_groups - 1) / p.tap_groups;
  }
  p.nt = pl.NT;
  pl.bias_fits = (pl.NT + 1) * p.Npad <= 512;             // room for the bias accumulator next to the tap accumulators
  int cols = (pl.NT + (pl.bias_fits ? 1 : 0)) * p.Npad;
  // nine taps per MMA (see Params::n3): flag 1024 keeps the three-accumulator form (A/B, tools/wgrad_phases.py)
  p.n3_cols = 3 * p.nchunks * 8;
  {
    const bool m64_later = p.Keff <= 16 && !(g_swap & 32);         // M = 64 takes N % 8 == 0, M = 128 needs N % 16 == 0 per row count
    p.n3 = (p.a3 && p.n3_cols <= 256 && (m64_later || (p.nchunks & 1) == 0) && !(g_swap & 1024)) ? 1 : 0;
  }
  if (p.n3) {
    pl.bias_fits = true;
    cols = p.n3_cols + p.Npad;
  }
  p.tmem_cols = 32;
  while ((int)p.tmem_cols < cols) p.tmem_cols *= 2;
  // narrow stride-1 layers on wide images: 8 x 32 tiles halve the per-tile hand-shakes, which is what those launches
  // are bound by once the MMAs are few (tools/wgrad_phases.py)
  pl.tu = (pl.mode == MODE_S1 && p.Keff <= 16 && (OW % 32) == 0 && !(g_swap & 128)) ? 32 : 16;   // Keff = 32: 71 vs 66 us, not worth it
  const int TU = pl.tu;
  p.tiles_u = (OW + TU - 1) / TU;
  p.tiles_v = (OH + TV - 1) / TV;
  p.ntiles = B * p.tiles_u * p.tiles_v;
  const int SP = pl.mode == MODE_S2 ? Geo<MODE_S2>::SP : (TU == 32 ? Geo<MODE_S1, 32>::SP : Geo<MODE_S1>::SP);
  const int PPAD = pl.mode == MODE_S2 ? Geo<MODE_S2>::PPAD : (TU == 32 ? Geo<MODE_S1, 32>::PPAD : Geo<MODE_S1>::PPAD);
  const int NSUB = pl.mode == MODE_S2 ? 4 : 1;
  const int DY_PLANE = TV * TU * 16;
  p.dy_bytes = (uint32_t)(p.nchunks * DY_PLANE);
  p.stage_bytes = (uint32_t)((SP * C * 2 + 127) / 128 * 128);
  // an M = 128 MMA always walks 16 plane strides from its start address and an N-padded one a plane past dY:
  // keep those reads inside the allocation
  // (8 strides for the M = 64 form)
  // (planes per CTA, buffers, shared-memory budget): deep pipelines while two CTAs still share an SM, then whatever fits
  const int tries[8][3] = {{16, 6, 110}, {16, 4, 110}, {16, 3, 110}, {16, 2, 110}, {16, 3, 220}, {16, 2, 220}, {16, 1, 220}, {8, 1, 220}};
  bool fits = false;
  for (int i = 0; i < 8 && !fits; ++i) {
    p.kpl = p.Keff / 8 < tries[i][0] ? p.Keff / 8 : tries[i][0];
    p.nbuf = tries[i][1];
    p.mblocks = (p.Keff / 8 + p.kpl - 1) / p.kpl;
    p.sub_bytes = (uint32_t)((p.a3 ? 3 : 1) * p.kpl * PPAD * 16);
    p.a_bytes = (NSUB * p.sub_bytes + 127) / 128 * 128;
    p.off_a = HDR;
    p.off_dy = p.off_a + p.nbuf * p.a_bytes;
    p.off_stage = p.off_dy + p.nbuf * p.dy_bytes + DY_PLANE;
    uint32_t total = p.off_stage + p.nbuf * p.stage_bytes;
    const bool m64 = (p.a3 ? 3 : (p.s4 ? 4 : 1)) * p.kpl <= 8 && !(g_swap & 32);
    const uint32_t overread = (m64 ? 8u : 16u) * PPAD * 16u + 4096u;
    const uint32_t need = p.off_a + (p.nbuf - 1) * p.a_bytes + (NSUB - 1) * p.sub_bytes + overread;
    if (total < need) total = need;
    pl.smem = (int)total + 128;
    fits = pl.smem <= tries[i][2] * 1024;
  }
  if (!fits) return false;
  int per_sm = (227 * 1024) / (pl.smem + 1024);
  const int tmem_lim = 512 / (int)p.tmem_cols;
  if (per_sm > tmem_lim) per_sm = tmem_lim;
  if (per_sm > 3) per_sm = 3;
  const int groups = p.tap_groups * p.mblocks;
  int S = (148 * per_sm) / groups;
  if (S < 1) S = 1;
  if (S > p.ntiles) S = p.ntiles;
  p.S = S;
  pl.S = S;
  pl.grid = S * groups;
  p.swap = g_swap;
  // the A operand is read from shared memory at 128 rows x 32 B per MMA whatever the real channel count: half of that
  // (M = 64, N % 8 == 0) when the CTA's channel block is at most 64 wide.  Flag 32 forces M = 128 (experiments).
  p.m64 = ((p.a3 ? 3 : (p.s4 ? 4 : 1)) * p.kpl <= 8 && !(g_swap & 32)) ? 1 : 0;
  p.idesc = idesc_mn(p.m64 ? 64 : 128, p.Npad);
  for (int c = 1; c <= 3; ++c) p.idesc_rows[c - 1] = idesc_mn(p.m64 ? 64 : 128, c * p.nchunks * 8);
  p.idesc_bias = p.idesc;
  return true;
}

template <int MODE, int TU = 16, bool N3 = false>
static int launch_k(const Plan& pl, cudaStream_t stream) {
  auto kern = k_wgrad_tc<MODE, TU, N3>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 221 * 1024);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(wgrad_tc): %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  kern<<<pl.grid, THREADS, pl.smem, stream>>>(pl.p);
  count_launch();
  return check_launch("k_wgrad_tc");
}

}  // namespace wg

int* tc_status_word();

long long wgrad_tc_workspace(int B, int H, int W, int C, int OH, int OW, int N, int ksize, int stride, int pro) {
  wg::Plan pl;
  if (!wg::plan(pl, B, H, W, C, OH, OW, N, ksize, stride, pro)) return 0;
  return (long long)pl.S * pl.p.T * pl.p.Keff * N + (long long)pl.S * N;      // weight partials, then bias partials
}

// returns false when this shape has no tensor-core kernel (the caller then runs the CUDA-core one);
// on success *S_out partials of [T][Keff][N] floats are in `partial`
// want_bias: also reduce dY over pixels; *bias_partial_out then points at [S][N] partials inside `partial`'s buffer (or is
// null when the extra accumulator does not fit TMEM and the caller has to run the bias kernel)
bool wgrad_tc_try(const void* x, const void* dY, float* partial, int B, int H, int W, int C, int OH, int OW, int N, int ksize,
                  int stride, int pro, float keep_prob, unsigned long long seed, int* S_out, cudaStream_t stream, int& rc,
                  bool want_bias, float** bias_partial_out, const unsigned long long* seed_dev) {
  using namespace wg;
  rc = 0;
  Plan pl;
  if (!plan(pl, B, H, W, C, OH, OW, N, ksize, stride, pro)) return false;
  Params& p = pl.p;
  PFN_encodeTiled enc = get_encode();
  {
    const int bw = pl.mode == MODE_S2 ? Geo<MODE_S2>::SW_ : pl.tu + 2, bh = pl.mode == MODE_S2 ? Geo<MODE_S2>::SH_ : Geo<MODE_S1>::SH_;
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    const cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)bw, (cuuint32_t)bh, 1u};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    if (enc(&p.tm_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return false;
  }
  {
    cuuint64_t dims[5] = {8, (cuuint64_t)(N / 8), (cuuint64_t)OW, (cuuint64_t)OH, (cuuint64_t)B};
    cuuint64_t strides[4] = {16, (cuuint64_t)N * 2, (cuuint64_t)OW * N * 2, (cuuint64_t)OH * OW * N * 2};
    cuuint32_t box[5] = {8, 1, (cuuint32_t)pl.tu, (cuuint32_t)TV, 1};
    if (p.n3) {      // x before the chunk: the box lands as [row][chunk][x][8]
      dims[1] = (cuuint64_t)OW; dims[2] = (cuuint64_t)(N / 8);
      strides[0] = (cuuint64_t)N * 2; strides[1] = 16;
      box[1] = (cuuint32_t)pl.tu; box[2] = (cuuint32_t)(N / 8);
    }
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    if (enc(&p.tm_dy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(dY), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return false;
  }
  p.partial = partial;
  p.bias_partial = (want_bias && pl.bias_fits) ? partial + (size_t)pl.S * p.T * p.Keff * N : nullptr;
  if (bias_partial_out) *bias_partial_out = p.bias_partial;
  p.keep_prob = keep_prob;
  p.seed = seed;
  p.seed_dev = seed_dev;
  p.status = tc_status_word();
  if (!p.status) { rc = (int)cudaErrorMemoryAllocation; return true; }
  *S_out = pl.S;
  if (p.n3) rc = pl.tu == 32 ? launch_k<MODE_S1, 32, true>(pl, stream) : launch_k<MODE_S1, 16, true>(pl, stream);
  else if (pl.mode == MODE_S1 && pl.tu == 32) rc = launch_k<MODE_S1, 32>(pl, stream);
  else if (pl.mode == MODE_S1) rc = launch_k<MODE_S1>(pl, stream);
  else rc = launch_k<MODE_S2>(pl, stream);
  return true;
}

}  // namespace fn

extern "C" {
void fn_debug_set_wgrad_swap(int v) { fn::wg::g_swap = v; }
}
