// tcgen05 / TMEM implicit-GEMM convolutions for sm_100a: stride-1 3x3 / 1x1 ('S1'), stride-2 3x3
// ('S2') and 3x3 stride-2 transposed ('TCONV') share one kernel.
//
// One CTA computes 16 x (8*J) GEMM rows ("row pixels": output pixels for S1/S2, INPUT pixels for
// TCONV whose four output phases are four accumulators) for all N output channels:
//
//   gather     all 160 threads read the raw bf16 NHWC halo tile with 4 independent 16-byte loads in
//              flight each, apply the prologue nonlinearity (concat_elu & co, reference
//              flow_architecture.py:14-29) in registers and write the UMMA A operand to shared
//              memory as 8-channel planes   A[sub][q][pv][pu][8]  (q = k/8; 16 B per pixel; no swizzle)
//              so that a filter tap is a pure start-address offset of the shared-memory descriptor:
//              GEMM row r = 8*v+u reads pixel (v+dv, u+du) -> SBO = PU*16, LBO = plane stride.
//              S2 splits the input tile into its four (row,col)-parity sub-planes so that the
//              stride-2 taps are again unit-stride; TCONV needs a one-pixel halo before the tile only.
//              Neither the 2C-channel concat_elu tensor nor an im2col matrix ever exists in HBM.
//   weights    pre-packed [kstep][2][N][8] bf16 (fn_pack_conv_weights_tc); streamed by the TMA
//              engine (cp.async.bulk + mbarrier) through a ring of <= 3 chunks of <= 32 KB
//   mma        one thread issues tcgen05.mma (M=128, N, K=16) per (k-step, sub-tile); fp32
//              accumulators live in TMEM (J * phases * N columns); the nin skip (reference :136-142)
//              is just extra k-steps on a second A operand with the centre tap
//   epilogue   4 warps read their TMEM quadrant (tcgen05.ld), add bias, apply gate*sigmoid +
//              shortcut (avg-pool / front channel pad, reference :150-165) and store bf16 NHWC, 16 B/lane
//
// Reference semantics: tf.nn.conv2d 'SAME' (flow_architecture.py:57-68), tf.nn.conv2d_transpose
// 'SAME' (:70-87, phase taps SURVEY App. B.2), res_block (:129-165).
#include "common.cuh"
#include "ptx.cuh"

namespace fn {

constexpr int TC_THREADS = 160;      // 4 gather/epilogue warps + 1 MMA/TMA warp (which also gathers)
constexpr int TC_MAX_STAGES = 3;
constexpr int TC_CHUNK_BYTES = 32768;
constexpr int TC_SMEM_HEADER = 128 + 1024;  // mbarriers + TMEM base pointer, then bias as float[256]
constexpr int TC_GATHER_UNROLL = 4;
constexpr int TC_SMEM_LIMIT = 200 * 1024;

enum { MODE_S1 = 0, MODE_S2 = 1, MODE_TCONV = 2 };

// debug knobs (tests only)
static int g_tc_swap_lbo_sbo = 0;
static int g_tc_use_s1p = 1;            // 0: force the generic kernel (tests compare the two)
static int* g_tc_status_dev = nullptr;   // device word: 0 ok, else first failing wait code
static long long* g_tc_timing_dev = nullptr;   // debug phase stamps
static int g_tc_timing_ctas = 0;

struct TcParams {
  const __nv_bfloat16* x;
  const __nv_bfloat16* skip;
  const __nv_bfloat16* wpk;
  const float* bias;
  void* y;
  const __nv_bfloat16* res;
  int* status;
  int mode;
  int B, H, W, C;            // input tensor
  int OH, OW;                // output tensor spatial size
  int Cs, SH, SW;            // skip tensor
  int pro, epi;
  int N, F, Cout;            // MMA N (multiple of 16); channels written per output pixel; real accumulator width
  int Cres, RH, RW, res_pool;
  int transposed, J;
  int EU, EV, tiles_u, tiles_v;   // GEMM-row image extents along the fast (u) / slow (v) axis
  int PU, P, Ppad;           // sub-plane row pitch, pixels, padded plane stride (pixels)
  uint32_t magic_P, magic_PU;
  int nsub;                  // sub-planes per operand (4 for S2, else 1)
  int chunk_log2;            // log2(C/8)
  int planes;                // Keff/8
  int org_u, org_v;          // sub-plane origin relative to the tile (halo before), S1/TCONV: 1, S2: pad
  int ntaps, nph;            // taps (9 or 1); accumulator phases (4 for TCONV, else 1)
  uint32_t tapoff[9];        // byte offset of each tap's A view relative to the operand base
  int tapcol[9];             // accumulator phase of each tap
  int tapfirst;              // bitmask: tap is the first one written into its phase
  uint32_t centre;           // byte offset of the centre view (skip operand)
  int ksteps_tap, ksteps_skip, ksteps_total;
  int half_k;
  int spc, nchunks, nstages;
  int a_main_bytes, a_skip_bytes;
  int tmem_cols;
  int swap_desc;
  int NI, split_phase;       // issuer warps; 1: issuers split by accumulator phase (TCONV) instead of by sub-tile
  long long* timing;         // debug: 8 clock64 stamps per CTA (nullptr in production)
  int timing_ctas;
};

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, int swap) {
  return swap ? ptx::smem_desc(addr, sbo, lbo) : ptx::smem_desc(addr, lbo, sbo);
}
__device__ __forceinline__ uint32_t fastdiv(uint32_t x, uint32_t magic) { return __umulhi(x, magic); }

// Gather one operand into A planes.  Item order: (sub-plane, pixel, chunk) with the chunk fastest so
// that a warp reads consecutive 16-byte pieces of consecutive pixels (coalesced); the padded plane
// stride keeps the 16-byte shared-memory stores of 8 neighbouring lanes on distinct banks.
__device__ __forceinline__ void gather_operand(const TcParams& p, const __nv_bfloat16* __restrict__ src, int Craw,
                                               int SHh, int SWw, int mode, int nsub, int b, int u0, int v0,
                                               unsigned char* planes_base) {
  const int clog = p.chunk_log2;          // same chunk count for skip and main in every supported layer
  const int chunks = Craw >> 3;
  const int items = nsub * p.P * chunks;
  const int mult = pro_mult(p.pro);
  const uint32_t sub_bytes = (uint32_t)(mult * chunks) * p.Ppad * 16u;
  for (int base = threadIdx.x; base < items; base += TC_THREADS * TC_GATHER_UNROLL) {
    uint4 raw[TC_GATHER_UNROLL];
    int dst[TC_GATHER_UNROLL];
#pragma unroll
    for (int k = 0; k < TC_GATHER_UNROLL; ++k) {
      const int idx = base + k * TC_THREADS;
      raw[k] = make_uint4(0u, 0u, 0u, 0u);
      dst[k] = -1;
      if (idx < items) {
        const int ch = idx & (chunks - 1);
        const uint32_t rest = (uint32_t)idx >> clog;
        const uint32_t sp = fastdiv(rest, p.magic_P);
        const uint32_t pp = rest - sp * p.P;
        const uint32_t pv = fastdiv(pp, p.magic_PU);
        const uint32_t pu = pp - pv * p.PU;
        int cu = u0 + (int)pu - p.org_u, cv = v0 + (int)pv - p.org_v;
        if (mode == MODE_S2) {              // parity sub-plane sp = 2*par_v + par_u holds input 2*pos + par
          cu = 2 * cu + (int)(sp & 1);
          cv = 2 * cv + (int)(sp >> 1);
        }
        const int yy = p.transposed ? cu : cv;
        const int xx = p.transposed ? cv : cu;
        if (cu >= 0 && cv >= 0 && yy < SHh && xx < SWw)
          raw[k] = __ldg(reinterpret_cast<const uint4*>(src + (((size_t)b * SHh + yy) * SWw + xx) * Craw + ch * 8));
        dst[k] = (int)(sp * sub_bytes + ((uint32_t)ch * p.Ppad + pp) * 16u);
      }
    }
#pragma unroll
    for (int k = 0; k < TC_GATHER_UNROLL; ++k) {
      if (dst[k] < 0) continue;
      bf16x8 in, o0, o1;
      in.v = raw[k];
      if (p.pro == FN_PRO_NONE) {
        o0.v = in.v;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float a0, a1;
          pro_apply(p.pro, __bfloat162float(in.h[j]), a0, a1);
          o0.h[j] = __float2bfloat16(a0);
          o1.h[j] = __float2bfloat16(a1);
        }
      }
      *reinterpret_cast<uint4*>(planes_base + dst[k]) = o0.v;
      if (mult == 2) *reinterpret_cast<uint4*>(planes_base + dst[k] + (uint32_t)chunks * p.Ppad * 16u) = o1.v;
    }
  }
}

// shortcut(res)[pixel, c0..c0+8) added into v[] (avg-pool and front channel pad, reference :153-165)
__device__ __forceinline__ void add_shortcut8(const TcParams& p, int b, int oy, int ox, int c0, float (&v)[8]) {
  if (p.res == nullptr) return;
  const int shift = p.F - p.Cres;
  if (!p.res_pool && shift == 0) {
    bf16x8 r;
    r.v = __ldg(reinterpret_cast<const uint4*>(p.res + (((size_t)b * p.RH + oy) * p.RW + ox) * p.Cres + c0));
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += __bfloat162float(r.h[j]);
    return;
  }
  if (p.res_pool && (p.Cres & 7) == 0 && ((c0 - shift) & 7) == 0 && 2 * oy + 1 < p.RH && 2 * ox + 1 < p.RW) {
    if (c0 - shift < 0) return;             // front padding: nothing to add
    const __nv_bfloat16* r00 = p.res + (((size_t)b * p.RH + 2 * oy) * p.RW + 2 * ox) * p.Cres + (c0 - shift);
    bf16x8 q[4];
    q[0].v = __ldg(reinterpret_cast<const uint4*>(r00));
    q[1].v = __ldg(reinterpret_cast<const uint4*>(r00 + p.Cres));
    q[2].v = __ldg(reinterpret_cast<const uint4*>(r00 + (size_t)p.RW * p.Cres));
    q[3].v = __ldg(reinterpret_cast<const uint4*>(r00 + (size_t)p.RW * p.Cres + p.Cres));
#pragma unroll
    for (int j = 0; j < 8; ++j)
      v[j] += 0.25f * (__bfloat162float(q[0].h[j]) + __bfloat162float(q[1].h[j]) + __bfloat162float(q[2].h[j]) +
                       __bfloat162float(q[3].h[j]));
    return;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int cr = c0 + j - shift;
    if (cr < 0) continue;
    if (p.res_pool) {
      float s = 0.f;
      int cnt = 0;
      for (int a = 0; a < 2; ++a)
        for (int c = 0; c < 2; ++c) {
          const int ry = 2 * oy + a, rx = 2 * ox + c;
          if (ry < p.RH && rx < p.RW) {
            s += __bfloat162float(p.res[(((size_t)b * p.RH + ry) * p.RW + rx) * p.Cres + cr]);
            ++cnt;
          }
        }
      v[j] += s / (float)cnt;
    } else {
      v[j] += __bfloat162float(p.res[(((size_t)b * p.RH + oy) * p.RW + ox) * p.Cres + cr]);
    }
  }
}

__global__ void __launch_bounds__(TC_THREADS, 6) k_conv_tc(const __grid_constant__ TcParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  // header: [0]=done, [1..3]=full[s], [4..6]=empty[s] (8 B each), byte 64 = TMEM base (4 B)
  const uint32_t smem_base = ptx::smem_u32(smem);
  const uint32_t bar_done = smem_base;
  const uint32_t bar_full = smem_base + 8;
  const uint32_t bar_empty = smem_base + 8 + 8 * TC_MAX_STAGES;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + 64);
  float* sbias = reinterpret_cast<float*>(smem + 128);
  unsigned char* a_main = smem + TC_SMEM_HEADER;
  unsigned char* a_skip = a_main + p.a_main_bytes;
  unsigned char* b_ring = a_skip + p.a_skip_bytes;
  const uint32_t b_ring_u32 = ptx::smem_u32(b_ring);
  const uint32_t chunk_bytes = (uint32_t)p.spc * p.N * 32u;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  long long* stamp = (p.timing != nullptr && (int)blockIdx.x < p.timing_ctas) ? p.timing + (size_t)blockIdx.x * 8 : nullptr;
  if (stamp && threadIdx.x == 0) stamp[0] = clock64();

  // ---- tile coordinates
  int t = blockIdx.x;
  const int tu = t % p.tiles_u;
  t /= p.tiles_u;
  const int tv = t % p.tiles_v;
  const int b = t / p.tiles_v;
  const int u0 = tu * 8 * p.J, v0 = tv * 16;

  // ---- setup: TMEM allocation + barriers + first weight chunks (warp 4)
  if (warp == 4) {
    ptx::tmem_alloc(smem_base + 64, (uint32_t)p.tmem_cols);
    ptx::tmem_relinquish();
    if (lane == 0) {
      ptx::mbar_init(bar_done, (uint32_t)p.NI);
      for (int s = 0; s < TC_MAX_STAGES; ++s) {
        ptx::mbar_init(bar_full + 8 * s, 1);
        ptx::mbar_init(bar_empty + 8 * s, (uint32_t)p.NI);
      }
      ptx::fence_mbar_init();
      // every ring slot starts empty: fill them all
      const int pre = p.nchunks < p.nstages ? p.nchunks : p.nstages;
      for (int c = 0; c < pre; ++c) {
        const int steps = min(p.spc, p.ksteps_total - c * p.spc);
        const uint32_t bytes = (uint32_t)steps * p.N * 32u;
        ptx::mbar_arrive_expect_tx(bar_full + 8 * c, bytes);
        ptx::bulk_g2s(b_ring_u32 + c * chunk_bytes, reinterpret_cast<const unsigned char*>(p.wpk) + (size_t)c * chunk_bytes,
                      bytes, bar_full + 8 * c);
      }
    }
    __syncwarp();
  }

  if (stamp && threadIdx.x == 0) stamp[1] = clock64();   // setup issued (warp 0 did none of it)
  // ---- gather + prologue: raw NHWC -> A planes
  for (int i = threadIdx.x; i < p.N; i += TC_THREADS) sbias[i] = i < p.Cout ? __ldg(p.bias + i) : 0.f;
  gather_operand(p, p.x, p.C, p.H, p.W, p.mode, p.nsub, b, u0, v0, a_main);
  if (p.skip != nullptr) gather_operand(p, p.skip, p.Cs, p.SH, p.SW, MODE_S1, 1, b, u0, v0, a_skip);

  ptx::fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the tensor-core (async) proxy
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (stamp && threadIdx.x == 0) stamp[2] = clock64();   // gather done, CTA synced

  bool ok = true;
  if (warp == 4) {
    // ================= weight-ring producer (one elected thread) =================
    // chunk c lives in slot c % nstages; the first nstages chunks were issued during setup
    if (lane == 0) {
      int s = 0;
      uint32_t parity = 0;
      for (int c = p.nstages; c < p.nchunks; ++c) {
        ok = ptx::mbar_wait(bar_empty + 8 * s, parity);      // every issuer's MMAs on the old chunk are done
        if (!ok) { atomicCAS(p.status, 0, 1); break; }
        const int nsteps = min(p.spc, p.ksteps_total - c * p.spc);
        const uint32_t bytes = (uint32_t)nsteps * p.N * 32u;
        ptx::mbar_arrive_expect_tx(bar_full + 8 * s, bytes);
        ptx::bulk_g2s(b_ring_u32 + (uint32_t)s * chunk_bytes,
                      reinterpret_cast<const unsigned char*>(p.wpk) + (size_t)c * chunk_bytes, bytes, bar_full + 8 * s);
        if (++s == p.nstages) { s = 0; parity ^= 1u; }
      }
    }
    __syncwarp();
  } else {
    // ================= MMA issue: warps 0..NI-1, one elected lane each, own accumulators =================
    // The warp runs the loop convergently (descriptor arithmetic stays on the uniform datapath); only
    // the tcgen05 instructions are predicated on the elected lane.  Issuer w owns sub-tiles j = w, w+NI, ..
    // (split_phase: accumulator phase w of every sub-tile), so no two issuers ever touch the same
    // TMEM columns and the accumulation order per output is fixed -> bit-reproducible.
    if (warp < p.NI) {
      const uint32_t leader = ptx::elect_one();
      const uint32_t idesc = ptx::idesc_bf16_f32(128, p.N);
      const uint32_t a_lbo = p.half_k ? 0u : (uint32_t)p.Ppad * 16u, a_sbo = (uint32_t)p.PU * 16u;
      const uint64_t a_tmpl = ptx::smem_desc(ptx::smem_u32(a_main), a_lbo, a_sbo);
      const uint64_t a_skip_tmpl = ptx::smem_desc(ptx::smem_u32(a_skip) + p.centre, a_lbo, a_sbo);
      const uint64_t b_tmpl = ptx::smem_desc(b_ring_u32, (uint32_t)p.N * 16u, 128u);
      const uint32_t kk_step16 = (2u * a_lbo) >> 4;      // two 8-channel planes per k-step
      const uint32_t b_step16 = ((uint32_t)p.N * 32u) >> 4;
      const uint32_t chunk16 = chunk_bytes >> 4;
      const uint32_t jcols = (uint32_t)(p.nph * p.N);     // TMEM columns per sub-tile
      const int J = p.J, NI = p.NI, split_phase = p.split_phase;
      const int j_first = split_phase ? 0 : warp, j_step = split_phase ? 1 : NI;
      int tap = 0, kk = 0;
      int ksteps_tap = p.ksteps_tap;
      bool in_skip = false;
      uint64_t a_desc = a_tmpl + (uint64_t)(p.tapoff[0] >> 4);
      uint32_t dcol = (uint32_t)(p.tapcol[0] * p.N);
      uint32_t acc = 0u;                                  // tap 0 is always the first of its phase
      bool mine = !split_phase || p.tapcol[0] == warp;
      int s = 0;
      uint32_t parity = 0;
      int ks_left = p.ksteps_total;
      for (int c = 0; c < p.nchunks && ok; ++c) {
        ok = ptx::mbar_wait(bar_full + 8 * s, parity);
        if (!ok) { atomicCAS(p.status, 0, 2); break; }
        ptx::tc_fence_after_sync();
        const int steps = min(p.spc, ks_left);
        ks_left -= steps;
        uint64_t b_desc = b_tmpl + (uint64_t)((uint32_t)s * chunk16);
        for (int i = 0; i < steps; ++i) {
          if (mine) {
            for (int j = j_first; j < J; j += j_step) {
              if (leader) ptx::mma_bf16_ss(tmem_base + (uint32_t)j * jcols + dcol, a_desc + (uint64_t)(8u * (uint32_t)j), b_desc, idesc, acc);
            }
          }
          b_desc += b_step16;
          a_desc += kk_step16;
          acc = 1u;
          if (++kk == ksteps_tap) {       // next tap (or the skip operand after the last one)
            kk = 0;
            ++tap;
            if (tap < p.ntaps) {
              a_desc = a_tmpl + (uint64_t)(p.tapoff[tap] >> 4);
              dcol = (uint32_t)(p.tapcol[tap] * p.N);
              acc = ((p.tapfirst >> tap) & 1) ? 0u : 1u;
              mine = !split_phase || p.tapcol[tap] == warp;
            } else if (!in_skip) {
              in_skip = true;
              a_desc = a_skip_tmpl;
              dcol = 0u;
              mine = !split_phase || warp == 0;
              ksteps_tap = 1 << 30;
            }
          }
        }
        __syncwarp();
        if (leader) ptx::mma_commit(bar_empty + 8 * s);   // this issuer is done reading ring slot s
        if (++s == p.nstages) { s = 0; parity ^= 1u; }
      }
      if (leader) {
        ptx::mma_commit(bar_done);                        // all of this issuer's MMAs complete -> arrive
        if (stamp && warp == 0) stamp[3] = clock64();     // all MMAs issued
      }
      __syncwarp();
    }
    // ================= epilogue: TMEM -> registers -> bf16 NHWC =================
    const int r = threadIdx.x;            // accumulator row == TMEM lane
    const int v = r >> 3, u = r & 7;
    // the shortcut does not depend on the accumulators: fetch it while the MMAs run
    const bool sc_plain = p.res != nullptr && !p.res_pool && p.Cres == p.F &&
                          (p.epi == FN_EPI_GATE_RES || p.epi == FN_EPI_RES);
    const bool sc_regs = sc_plain && p.F == 8;     // the 128x256 level: one 16-byte piece per sub-tile, keep in registers
    uint4 sc_pre[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) sc_pre[j] = make_uint4(0u, 0u, 0u, 0u);
    if (sc_plain) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j >= p.J) break;
        const int cu = u0 + 8 * j + u, cv = v0 + v;
        if (cu < p.EU && cv < p.EV) {
          const int ry = p.transposed ? cu : cv, rx = p.transposed ? cv : cu;
          const __nv_bfloat16* rp = p.res + (((size_t)b * p.RH + ry) * p.RW + rx) * p.Cres;
          if (sc_regs) {
            sc_pre[j] = __ldg(reinterpret_cast<const uint4*>(rp));
          } else {
            for (int c0 = 0; c0 < p.F; c0 += 64) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + c0));
          }
        }
      }
    }
    ok = ptx::mbar_wait(bar_done, 0);
    if (!ok) atomicCAS(p.status, 0, 3);
    ptx::tc_fence_after_sync();
    if (stamp && threadIdx.x == 0) stamp[4] = clock64();   // accumulators ready
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    const uint32_t jcols = (uint32_t)(p.nph * p.N);
    __nv_bfloat16* yb = reinterpret_cast<__nv_bfloat16*>(p.y);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j >= p.J) break;
      const int cu = u0 + 8 * j + u, cv = v0 + v;
      const bool live = ok && cu < p.EU && cv < p.EV;
      const int ry = p.transposed ? cu : cv, rx = p.transposed ? cv : cu;    // GEMM-row pixel in image coords
      const uint32_t col0 = lane_base + (uint32_t)j * jcols;
      if (p.mode == MODE_TCONV) {
        for (int ph = 0; ph < 4; ++ph) {
          const size_t pix = ((size_t)b * p.OH + (2 * ry + (ph >> 1))) * p.OW + (2 * rx + (ph & 1));
          for (int c0 = 0; c0 < p.F; c0 += 8) {
            float o[8];
            ptx::tmem_ld8(col0 + (uint32_t)(ph * p.N + c0), o);
            if (live) {
              bf16x8 st;
#pragma unroll
              for (int i = 0; i < 8; ++i) st.h[i] = __float2bfloat16(o[i] + sbias[c0 + i]);
              *reinterpret_cast<uint4*>(yb + pix * p.F + c0) = st.v;
            }
          }
        }
      } else if (p.epi == FN_EPI_GATE_RES) {
        const size_t pix = ((size_t)b * p.OH + ry) * p.OW + rx;
        for (int c0 = 0; c0 < p.F; c0 += 8) {
          float a[8], g[8];
          ptx::tmem_ld8x2(col0 + c0, col0 + p.F + c0, a, g);
          if (live) {
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
              o[i] = (a[i] + sbias[c0 + i]) * sigmoid_f(g[i] + sbias[p.F + c0 + i]);
            if (sc_regs) {
              bf16x8 rr;
              rr.v = sc_pre[j];
#pragma unroll
              for (int i = 0; i < 8; ++i) o[i] += __bfloat162float(rr.h[i]);
            } else {
              add_shortcut8(p, b, ry, rx, c0, o);
            }
            bf16x8 st;
#pragma unroll
            for (int i = 0; i < 8; ++i) st.h[i] = __float2bfloat16(o[i]);
            *reinterpret_cast<uint4*>(yb + pix * p.F + c0) = st.v;
          }
        }
      } else {
        const size_t pix = ((size_t)b * p.OH + ry) * p.OW + rx;
        for (int c0 = 0; c0 < p.F; c0 += 8) {
          float o[8];
          ptx::tmem_ld8(col0 + c0, o);
          if (live) {
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] += sbias[c0 + i];
            if (p.epi == FN_EPI_RES) add_shortcut8(p, b, ry, rx, c0, o);
            bf16x8 st;
#pragma unroll
            for (int i = 0; i < 8; ++i) st.h[i] = __float2bfloat16(o[i]);
            *reinterpret_cast<uint4*>(yb + pix * p.F + c0) = st.v;
          }
        }
      }
    }
  }

  // ---- teardown
  if (stamp && threadIdx.x == 0) stamp[5] = clock64();     // epilogue of warp 0 done
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (stamp && threadIdx.x == 0) stamp[6] = clock64();
  if (warp == 4) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------ host side
static int round_up(int a, int m) { return (a + m - 1) / m * m; }
static int pow2_cols(int c) {
  int r = 32;
  while (r < c) r <<= 1;
  return r;
}
static int ilog2_exact(int v) {   // -1 unless v is a power of two
  int l = 0;
  while ((1 << l) < v) ++l;
  return (1 << l) == v ? l : -1;
}
static uint32_t magic_of(int d) { return (uint32_t)((1ull << 32) / (uint32_t)d + 1ull); }

// plane stride (pixels) such that 8 neighbouring gather lanes store to distinct 16-byte bank groups
static int padded_plane(int P, int chunks) {
  if (chunks == 1) return P;
  const int want = chunks >= 8 ? 1 : (chunks == 4 ? 2 : 4);
  int Pp = P;
  while ((Pp & 7) != want) ++Pp;
  return Pp;
}

struct TcGeom {   // everything plan_tc() needs to know about one launch, independent of the descriptor type
  int mode, B, H, W, C, OH, OW, Cout, ksize, pro, epi;
  int Cs, SH, SW;
  int pad_y, pad_x;   // S2: TF SAME pad-before per image axis
};

static bool plan_tc(const TcGeom& g, TcParams& p, size_t& smem_bytes, const char** why) {
  const char* dummy;
  if (!why) why = &dummy;
  const int mult = pro_mult(g.pro);
  const int Keff = g.C * mult;
  const int Kskip = g.Cs * mult;
  // Keff = 8 (8-channel gradients in the backward pass): one k-step that reads the single plane for both K halves
  // (LBO = 0) against weights whose second half is zero
  const bool half_k = Keff == 8 && Kskip == 0;
  if ((g.C % 8) != 0 || ((Keff % 16) != 0 && !half_k) || (Kskip % 16) != 0) { *why = "K per tap not a multiple of 16"; return false; }
  const int clog = ilog2_exact(g.C / 8);
  if (clog < 0 || (g.Cs && g.Cs != g.C)) { *why = "channel count not 8 * 2^n (or skip width differs)"; return false; }
  const int N = round_up(g.Cout, 16);
  if (N < 16 || N > 256) { *why = "N out of range"; return false; }
  const bool gated = g.epi == FN_EPI_GATE_RES;

  p.mode = g.mode;
  p.B = g.B; p.H = g.H; p.W = g.W; p.C = g.C; p.OH = g.OH; p.OW = g.OW;
  p.Cs = g.Cs; p.SH = g.SH; p.SW = g.SW;
  p.pro = g.pro; p.epi = g.epi;
  p.N = N; p.F = gated ? g.Cout / 2 : g.Cout; p.Cout = g.Cout;
  p.ntaps = g.ksize * g.ksize;
  p.nph = g.mode == MODE_TCONV ? 4 : 1;
  p.nsub = g.mode == MODE_S2 ? 4 : 1;
  p.chunk_log2 = clog;
  p.planes = Keff / 8;
  p.ksteps_tap = (Keff + 15) / 16;
  p.half_k = half_k ? 1 : 0;
  p.ksteps_skip = Kskip / 16;
  p.ksteps_total = p.ntaps * p.ksteps_tap + p.ksteps_skip;
  auto set_ring = [&](int chunk_budget, int max_stages) {
    p.spc = chunk_budget / (N * 32);
    if (p.spc < 1) p.spc = 1;
    if (p.spc > p.ksteps_total) p.spc = p.ksteps_total;
    p.nchunks = (p.ksteps_total + p.spc - 1) / p.spc;
    p.nstages = p.nchunks < max_stages ? p.nchunks : max_stages;
  };
  p.swap_desc = g_tc_swap_lbo_sbo;
  p.timing = g_tc_timing_dev;
  p.timing_ctas = g_tc_timing_ctas;

  // GEMM-row image: output pixels (S1/S2) or input pixels (TCONV)
  const int RHh = g.mode == MODE_TCONV ? g.H : g.OH, RWw = g.mode == MODE_TCONV ? g.W : g.OW;
  // orientation: the 8-wide fast axis u runs along x unless the row image is too short for 16 rows
  p.transposed = (RHh < 16 && RWw >= 16) ? 1 : 0;
  p.EU = p.transposed ? RHh : RWw;
  p.EV = p.transposed ? RWw : RHh;
  p.tiles_v = (p.EV + 15) / 16;
  const int pad_u = p.transposed ? g.pad_y : g.pad_x, pad_v = p.transposed ? g.pad_x : g.pad_y;

  const int halo = g.mode == MODE_S1 ? 2 : 1;        // extra sub-plane rows/cols beyond the tile
  const int PV = 16 + halo;
  const long long want_ctas = 2LL * 148;
  int bestJ = 0;
  size_t best_sm = 0;
  // weight ring: prefer 3 x 32 KB; shrink it before giving up on a tile that would otherwise not fit
  const int ring_opts[3][2] = {{TC_CHUNK_BYTES, 3}, {TC_CHUNK_BYTES, 2}, {TC_CHUNK_BYTES / 2, 2}};
  for (int ro = 0; ro < 3 && bestJ == 0; ++ro) {
    set_ring(ring_opts[ro][0], ring_opts[ro][1]);
    for (int J = 4; J >= 1; J >>= 1) {   // largest J that still gives two waves of CTAs, else the smallest feasible
      if (J * p.nph * N > 512) continue;
      if (J * p.nph * N > 256 && J > 1) continue;
      if (J > 1 && (p.EU % (8 * J)) != 0) continue;
      const int PU = 8 * J + halo, P = PV * PU, Pp = padded_plane(P, g.C / 8);
      const size_t sm = TC_SMEM_HEADER + (size_t)p.nsub * p.planes * Pp * 16 + (size_t)(Kskip / 8) * Pp * 16 +
                        (size_t)p.nstages * p.spc * N * 32;
      if (sm > (size_t)TC_SMEM_LIMIT) continue;
      bestJ = J;
      best_sm = sm;
      const long long ctas = (long long)g.B * p.tiles_v * ((p.EU + 8 * J - 1) / (8 * J));
      if (ctas >= want_ctas) break;
    }
  }
  if (bestJ == 0) { *why = "tile does not fit shared memory / TMEM"; return false; }
  p.J = bestJ;
  p.PU = 8 * p.J + halo;
  p.P = PV * p.PU;
  p.Ppad = padded_plane(p.P, g.C / 8);
  p.magic_P = magic_of(p.P);
  p.magic_PU = magic_of(p.PU);
  p.tiles_u = (p.EU + 8 * p.J - 1) / (8 * p.J);
  p.a_main_bytes = p.nsub * p.planes * p.Ppad * 16;
  p.a_skip_bytes = (Kskip / 8) * p.Ppad * 16;
  p.tmem_cols = pow2_cols(p.J * p.nph * N);
  smem_bytes = best_sm;
  p.split_phase = (p.nph == 4 && p.J < 4) ? 1 : 0;
  p.NI = p.split_phase ? 4 : p.J;

  // ---- tap program
  p.tapfirst = 0;
  p.centre = (uint32_t)(p.PU + 1) * 16u;
  const uint32_t sub_bytes = (uint32_t)p.planes * p.Ppad * 16u;
  if (g.mode == MODE_S1) {
    p.org_u = p.org_v = 1;
    if (p.ntaps == 1) {
      p.tapoff[0] = p.centre;
      p.tapcol[0] = 0;
    } else {
      for (int t = 0; t < 9; ++t) {
        const int di = t / 3, dj = t % 3;
        const int dv = p.transposed ? dj : di, du = p.transposed ? di : dj;
        p.tapoff[t] = (uint32_t)(dv * p.PU + du) * 16u;
        p.tapcol[t] = 0;
      }
    }
    p.tapfirst = 1;
  } else if (g.mode == MODE_S2) {
    // input coordinate = 2*o + e, e = d - pad in {-1,0,1,2}: parity e&1, position o + floor(e/2);
    // the sub-plane origin sits `pad` positions before the tile
    p.org_u = pad_u;
    p.org_v = pad_v;
    for (int t = 0; t < 9; ++t) {
      const int di = t / 3, dj = t % 3;
      const int dv = p.transposed ? dj : di, du = p.transposed ? di : dj;
      const int ev = dv - pad_v, eu = du - pad_u;
      const int parv = ev & 1, paru = eu & 1;
      const int lv = (ev >= 0 ? ev / 2 : -1) + pad_v, lu = (eu >= 0 ? eu / 2 : -1) + pad_u;
      p.tapoff[t] = (uint32_t)(parv * 2 + paru) * sub_bytes + (uint32_t)(lv * p.PU + lu) * 16u;
      p.tapcol[t] = 0;
    }
    p.tapfirst = 1;
  } else {   // MODE_TCONV: out[2m+d-?]: taps d=0,1 read in[m], d=2 reads in[m-1]; phase = (d==1)
    p.org_u = p.org_v = 1;
    int seen = 0;
    for (int t = 0; t < 9; ++t) {
      const int di = t / 3, dj = t % 3;
      const int py = (di == 1), px = (dj == 1);
      const int sy = (di == 2) ? 0 : 1, sx = (dj == 2) ? 0 : 1;     // sub-plane position of the row pixel's source
      const int sv = p.transposed ? sx : sy, su = p.transposed ? sy : sx;
      p.tapoff[t] = (uint32_t)(sv * p.PU + su) * 16u;
      p.tapcol[t] = py * 2 + px;
      if (!((seen >> p.tapcol[t]) & 1)) {
        p.tapfirst |= 1 << t;
        seen |= 1 << p.tapcol[t];
      }
    }
  }
  return true;
}

static bool geom_of_conv(const fn_conv_desc& d, TcGeom& g, const char** why) {
  const char* dummy;
  if (!why) why = &dummy;
  if (d.epilogue == FN_EPI_TANH) { *why = "tanh tail runs on the CUDA-core kernel"; return false; }
  if (d.keep_prob < 1.0f) { *why = "dropout prologue"; return false; }
  if (d.w_tc == nullptr) { *why = "w_tc is null"; return false; }
  if (d.stride == 2 && d.ksize != 3) { *why = "stride 2 needs a 3x3 kernel"; return false; }
  if (d.stride == 2 && d.skip) { *why = "stride 2 with a skip operand"; return false; }
  memset(&g, 0, sizeof(g));
  g.mode = d.stride == 2 ? MODE_S2 : MODE_S1;
  g.B = d.B; g.H = d.H; g.W = d.W; g.C = d.Cin;
  g.OH = (d.H + d.stride - 1) / d.stride;
  g.OW = (d.W + d.stride - 1) / d.stride;
  g.Cout = d.Cout; g.ksize = d.ksize; g.pro = d.prologue; g.epi = d.epilogue;
  if (d.skip) { g.Cs = d.Cskip; g.SH = d.SH; g.SW = d.SW; }
  g.pad_y = same_pad_before(d.H, d.ksize, d.stride);
  g.pad_x = same_pad_before(d.W, d.ksize, d.stride);
  return true;
}

bool conv_s1p_try(const fn_conv_desc& d, cudaStream_t stream, bool dry_run, int& rc);   // conv_s1p.cu
bool conv_s1d_try(const fn_conv_desc& d, cudaStream_t stream, bool dry_run, int& rc);   // conv_s1d.cu
bool conv_s1n_try(const fn_conv_desc& d, cudaStream_t stream, bool dry_run, int& rc);   // conv_s2t.cu
bool conv_s1w_try(const fn_conv_desc& d, cudaStream_t stream, bool dry_run, int& rc);   // conv_s1w.cu

bool conv_pl_try(const fn_conv_desc& d, cudaStream_t stream, bool dry_run, int& rc);    // conv_pl.cu
bool conv_s2p_try(const fn_conv_desc& d, cudaStream_t stream, bool dry_run, int& rc);   // conv_s2t.cu

bool conv_tc_supported(const fn_conv_desc& d, const char** why) {
  int rc = 0;
  if (d.x_planes || d.skip_planes) {       // MMA-ready operands: the planes kernel or nothing
    if (conv_pl_try(d, nullptr, true, rc)) return true;
    if (why) *why = "planes operands are taken by stride-1 3x3 concat_elu convs with 8/16/32 channels only (conv_pl.cu)";
    return false;
  }
  if (d.y_planes || d.y == nullptr) {      // planes output from raw input: the persistent stride-2 kernel (the head runs on CUDA cores)
    if (g_tc_use_s1p && conv_s2p_try(d, nullptr, true, rc)) return true;
    if (why) *why = "a planes output is written by conv_pl.cu, the persistent stride-2 kernel and the 1-channel head only";
    return false;
  }
  // the persistent kernels also take the dropout prologue, which the generic kernel below does not
  if (g_tc_use_s1p && (d.keep_prob < 1.0f || d.epilogue == FN_EPI_GATE_BWD) &&
      (conv_s1p_try(d, nullptr, true, rc) || conv_s1d_try(d, nullptr, true, rc)))
    return true;
  if (g_tc_use_s1p && conv_s1w_try(d, nullptr, true, rc)) return true;      // C = 256: two-pass kernel, pass-major weights
  if (d.epilogue == FN_EPI_GATE_BWD) {
    if (why) *why = "the gate-backward epilogue exists in the persistent kernels only";
    return false;
  }
  if (d.epilogue == FN_EPI_PRO_BWD) {
    if (g_tc_use_s1p && conv_s1n_try(d, nullptr, true, rc)) return true;
    if (why) *why = "the prologue-adjoint epilogue exists in the persistent data-gradient kernel only";
    return false;
  }
  TcGeom g;
  TcParams p;
  size_t sm;
  memset(&p, 0, sizeof(p));
  return geom_of_conv(d, g, why) && plan_tc(g, p, sm, why);
}

static int ensure_status_word() {
  if (g_tc_status_dev == nullptr) {
    cudaError_t e = cudaMalloc(&g_tc_status_dev, sizeof(int));
    if (e != cudaSuccess) { set_error("cudaMalloc(status): %s", cudaGetErrorString(e)); return (int)e; }
    e = cudaMemset(g_tc_status_dev, 0, sizeof(int));
    if (e != cudaSuccess) { set_error("cudaMemset(status): %s", cudaGetErrorString(e)); return (int)e; }
  }
  return 0;
}

void tc_timing_buffer(long long** buf, int* ctas) { *buf = g_tc_timing_dev; *ctas = g_tc_timing_ctas; }
int* tc_status_word() { return ensure_status_word() == 0 ? g_tc_status_dev : nullptr; }
bool conv_s1p_try(const fn_conv_desc& d, cudaStream_t stream, bool dry_run, int& rc);   // conv_s1p.cu
bool conv_s1d_try(const fn_conv_desc& d, cudaStream_t stream, bool dry_run, int& rc);   // conv_s1d.cu
bool conv_s2p_try(const fn_conv_desc& d, cudaStream_t stream, bool dry_run, int& rc);   // conv_s2t.cu
bool conv_s1n_try(const fn_conv_desc& d, cudaStream_t stream, bool dry_run, int& rc);   // conv_s2t.cu
bool conv_s2n_try(const fn_conv_desc& d, cudaStream_t stream, bool dry_run, int& rc);   // conv_s2t.cu
bool tconv_s2p_try(const fn_tconv_desc& d, cudaStream_t stream, bool dry_run, int& rc); // conv_s2t.cu

static int launch_tc(TcParams& p, size_t smem_bytes, cudaStream_t stream) {
  int rc = ensure_status_word();
  if (rc) return rc;
  p.status = g_tc_status_dev;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_conv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  const unsigned grid = (unsigned)((long long)p.B * p.tiles_u * p.tiles_v);
  k_conv_tc<<<grid, TC_THREADS, smem_bytes, stream>>>(p);
  count_launch();
  return check_launch("k_conv_tc");
}

bool conv_wg_try(const fn_conv_desc& d, cudaStream_t stream, bool dry_run, int& rc);    // conv_wg.cu

// true if the kernel conv_fwd_tc() will pick for d reads the pass-major weight image (k_conv_s1w without a skip operand);
// follows the dispatch order below
bool conv_weights_pass_major(const fn_conv_desc& d) {
  int rc = 0;
  if (!g_tc_use_s1p || d.x_planes || d.skip_planes || d.y_planes || d.y == nullptr || d.skip != nullptr) return false;
  if (conv_wg_try(d, nullptr, true, rc) || conv_s1p_try(d, nullptr, true, rc) || conv_s1d_try(d, nullptr, true, rc)) return false;
  return conv_s1w_try(d, nullptr, true, rc);
}

int conv_fwd_tc(const fn_conv_desc& d, cudaStream_t stream) {
  int rc_sp = 0;
  if (d.x_planes || d.skip_planes) {
    if (conv_pl_try(d, stream, false, rc_sp)) return rc_sp;
    set_error("conv_fwd_tc: no kernel for these planes operands");
    return FN_E_UNSUPPORTED;
  }
  if (d.y_planes || d.y == nullptr) {
    if (g_tc_use_s1p && conv_s2p_try(d, stream, false, rc_sp)) return rc_sp;
    set_error("conv_fwd_tc: no kernel writes a planes output for this shape");
    return FN_E_UNSUPPORTED;
  }
  if (g_tc_use_s1p && conv_wg_try(d, stream, false, rc_sp)) return rc_sp;
  if (d.res_u8) {
    set_error("conv_fwd_tc: only k_conv_wg reads a uint8 shortcut, and it does not take this shape");
    return FN_E_UNSUPPORTED;
  }
  if (g_tc_use_s1p && conv_s1p_try(d, stream, false, rc_sp)) return rc_sp;
  if (g_tc_use_s1p && conv_s1d_try(d, stream, false, rc_sp)) return rc_sp;
  if (g_tc_use_s1p && conv_s1w_try(d, stream, false, rc_sp)) return rc_sp;
  if (g_tc_use_s1p && conv_s2p_try(d, stream, false, rc_sp)) return rc_sp;
  if (g_tc_use_s1p && conv_s1n_try(d, stream, false, rc_sp)) return rc_sp;
  if (g_tc_use_s1p && conv_s2n_try(d, stream, false, rc_sp)) return rc_sp;
  TcGeom g;
  TcParams p;
  size_t smem_bytes = 0;
  const char* why = "";
  memset(&p, 0, sizeof(p));
  if (!geom_of_conv(d, g, &why) || !plan_tc(g, p, smem_bytes, &why)) {
    set_error("conv_fwd_tc: unsupported: %s", why);
    return FN_E_UNSUPPORTED;
  }
  p.x = (const __nv_bfloat16*)d.x;
  p.skip = (const __nv_bfloat16*)d.skip;
  p.wpk = (const __nv_bfloat16*)d.w_tc;
  p.bias = d.bias;
  p.y = d.y;
  p.res = (const __nv_bfloat16*)d.res;
  p.Cres = d.Cres; p.RH = d.RH; p.RW = d.RW; p.res_pool = d.res_pool;
  return launch_tc(p, smem_bytes, stream);
}

static void geom_of_tconv(const fn_tconv_desc& d, TcGeom& g) {
  memset(&g, 0, sizeof(g));
  g.mode = MODE_TCONV;
  g.B = d.B; g.H = d.H; g.W = d.W; g.C = d.Cin; g.OH = 2 * d.H; g.OW = 2 * d.W;
  g.Cout = d.Cout; g.ksize = 3; g.pro = FN_PRO_NONE; g.epi = FN_EPI_BIAS;
}

bool tconv_tc_supported(const fn_tconv_desc& d, const char** why) {
  if (d.w_tc == nullptr) {
    if (why) *why = "w_tc is null";
    return false;
  }
  if (d.epilogue == FN_EPI_PRO_BWD) {
    int rc = 0;
    if (g_tc_use_s1p && tconv_s2p_try(d, nullptr, true, rc)) return true;
    if (why) *why = "the prologue-adjoint epilogue exists in the persistent sub-pixel kernel only";
    return false;
  }
  if (d.y_planes) {
    int rc = 0;
    if (g_tc_use_s1p && tconv_s2p_try(d, nullptr, true, rc)) return true;
    if (why) *why = "a planes output is written by the persistent sub-pixel kernel only";
    return false;
  }
  TcGeom g;
  TcParams p;
  size_t sm;
  memset(&p, 0, sizeof(p));
  geom_of_tconv(d, g);
  return plan_tc(g, p, sm, why);
}

int tconv_fwd_tc(const fn_tconv_desc& d, cudaStream_t stream) {
  int rc_sp = 0;
  if (g_tc_use_s1p && tconv_s2p_try(d, stream, false, rc_sp)) return rc_sp;
  if (d.y_planes) {
    set_error("tconv_fwd_tc: no kernel writes a planes output for this shape");
    return FN_E_UNSUPPORTED;
  }
  TcGeom g;
  TcParams p;
  size_t smem_bytes = 0;
  const char* why = "";
  memset(&p, 0, sizeof(p));
  geom_of_tconv(d, g);
  if (!plan_tc(g, p, smem_bytes, &why)) {
    set_error("tconv_fwd_tc: unsupported: %s", why);
    return FN_E_UNSUPPORTED;
  }
  p.x = (const __nv_bfloat16*)d.x;
  p.wpk = (const __nv_bfloat16*)d.w_tc;
  p.bias = d.bias;
  p.y = d.y;
  return launch_tc(p, smem_bytes, stream);
}

// ---- weight packing: src [T][Keff][Cout] (+ [Kskip][Cout]) bf16 -> dst [kstep][2][N][8] bf16
__global__ void k_pack_tc(const __nv_bfloat16* __restrict__ w, const __nv_bfloat16* __restrict__ ws,
                          __nv_bfloat16* __restrict__ dst, int T, int Keff, int Kskip, int Cout, int N) {
  const int kst = (Keff + 15) / 16;      // Keff = 8: one k-step whose second half is zero (the kernel reads the plane twice)
  const int main_steps = T * kst;
  const int total = (main_steps + Kskip / 16) * 2 * N * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int e = i & 7;
    int r = i >> 3;
    const int n = r % N;
    r /= N;
    const int h = r & 1;
    const int ks = r >> 1;
    __nv_bfloat16 v = __float2bfloat16(0.f);
    if (n < Cout) {
      if (ks < main_steps) {
        const int tap = ks / kst, kk = ks % kst;
        const int k = kk * 16 + h * 8 + e;
        if (k < Keff) v = w[((size_t)tap * Keff + k) * Cout + n];
      } else {
        const int k = (ks - main_steps) * 16 + h * 8 + e;
        v = ws[(size_t)k * Cout + n];
      }
    }
    dst[i] = v;
  }
}

// the same image straight from an fp32 variable [T][A][B] with the taps optionally reversed and the two channel axes
// optionally swapped -- the B operands of the data-gradient convs (stride-1 conv: flipped taps, K and N exchanged) without
// the flip / permute / cast / contiguous round trip through bf16 temporaries
__global__ void k_pack_tc_f32(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int T, int K, int Cout, int N, int flip,
                              int transpose) {
  const int kst = (K + 15) / 16;
  const int total = T * kst * 2 * N * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int e = i & 7;
    int r = i >> 3;
    const int n = r % N;
    r /= N;
    const int h = r & 1;
    const int ks = r >> 1;
    const int tap = ks / kst, kk = ks % kst;
    const int k = kk * 16 + h * 8 + e;
    float v = 0.f;
    if (n < Cout && k < K) {
      const int st = flip ? T - 1 - tap : tap;
      v = transpose ? w[((size_t)st * Cout + n) * K + k] : w[((size_t)st * K + k) * Cout + n];
    }
    dst[i] = __float2bfloat16(v);
  }
}

int pack_conv_weights_tc_f32(const float* w, void* dst, int ksize, int K, int Cout, int flip, int transpose, cudaStream_t stream) {
  FN_REQUIRE((K % 16) == 0 || K == 8, FN_E_ALIGN, "pack_conv_weights_tc_f32: K = %d must be a multiple of 16 (or 8)", K);
  const int N = round_up(Cout, 16);
  const long long total = (long long)ksize * ksize * ((K + 15) / 16) * 2 * N * 8;
  k_pack_tc_f32<<<(int)((total + 255) / 256), 256, 0, stream>>>(w, (__nv_bfloat16*)dst, ksize * ksize, K, Cout, N, flip, transpose);
  count_launch();
  return check_launch("k_pack_tc_f32");
}

long long conv_weights_tc_bytes(int ksize, int Keff, int Kskip, int Cout, int gated) {
  (void)gated;
  if (((Keff % 16) != 0 && Keff != 8) || (Kskip % 16) != 0 || Cout <= 0) return 0;
  const long long steps = (long long)ksize * ksize * ((Keff + 15) / 16) + Kskip / 16;
  return steps * round_up(Cout, 16) * 32;
}

int pack_conv_weights_tc(const void* w, const void* w_skip, void* dst, int ksize, int Keff, int Kskip, int Cout,
                         int gated, cudaStream_t stream) {
  (void)gated;
  FN_REQUIRE(((Keff % 16) == 0 || Keff == 8) && (Kskip % 16) == 0, FN_E_ALIGN,
             "pack_conv_weights_tc: K (%d,%d) must be multiples of 16 (or K = 8 without a skip operand)", Keff, Kskip);
  FN_REQUIRE(Kskip == 0 || w_skip != nullptr, FN_E_INVAL, "pack_conv_weights_tc: null w_skip");
  const int N = round_up(Cout, 16);
  const long long total = conv_weights_tc_bytes(ksize, Keff, Kskip, Cout, gated) / 2;
  const int grid = (int)((total + 255) / 256);
  k_pack_tc<<<grid, 256, 0, stream>>>((const __nv_bfloat16*)w, (const __nv_bfloat16*)w_skip, (__nv_bfloat16*)dst,
                                      ksize * ksize, Keff, Kskip, Cout, N);
  count_launch();
  return check_launch("k_pack_tc");
}

}  // namespace fn

// ---- debug / test hooks (not part of the drop-in surface, but exported for tests)
extern "C" {
void fn_debug_set_swap_lbo_sbo(int v) { fn::g_tc_swap_lbo_sbo = v; }
void fn_debug_set_use_s1p(int v) { fn::g_tc_use_s1p = v; }
// register (or clear with nullptr) a device buffer of 8 int64 clock stamps per CTA for the first `ctas` CTAs
void fn_debug_set_timing(long long* dev_buf, int ctas) {
  fn::g_tc_timing_dev = dev_buf;
  fn::g_tc_timing_ctas = ctas;
}
// synchronises the device; returns the tcgen05 protocol status word (0 = ok) and clears it
int fn_debug_tc_status(void) {
  if (fn::g_tc_status_dev == nullptr) return 0;
  int h = 0;
  cudaDeviceSynchronize();
  cudaMemcpy(&h, fn::g_tc_status_dev, sizeof(int), cudaMemcpyDeviceToHost);
  cudaMemset(fn::g_tc_status_dev, 0, sizeof(int));
  return h;
}
}
