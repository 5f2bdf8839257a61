// Warpgroup-autonomous persistent tcgen05 kernel for the stride-1 3x3 convolutions of flow_net with the concat_elu
// prologue and resident weights (C in {8,16,32}: the 128x256, 64x128 and 32x64 levels) -- round-2 replacement of
// k_conv_s1p on the inference path.
//
// k_conv_s1p hands every tile through four warp roles (TMA -> front -> issuer -> back); its knock-out profile (DESIGN.md
// section 3) showed the stage latencies ADD per tile and an empty pipeline already costs 8-16 us per launch.  Here a CTA
// is NWG independent warpgroups of 128 threads that share only the resident weights and one TMEM allocation.  Each
// warpgroup owns a stage buffer, a set of A planes and J accumulators, and walks its own tiles through the whole chain:
//
//   wait TMA(tile)  ->  all 128 threads: stage -> concat_elu -> UMMA A planes  ->  named barrier (128 threads)
//   -> one elected lane: TMA(next tile) into the now free stage, then the unrolled tcgen05.mma program, one commit per
//      128-pixel sub-tile  ->  all 128 threads: shortcut loads, wait commit j, TMEM -> bias / gate / shortcut -> bf16 NHWC
//
// i.e. two light hand-offs per tile (one bar.sync, one mbarrier) instead of five, and the overlap comes from 6-8
// warpgroups per SM being in different phases -- latency hiding by occupancy, as for any CUDA kernel, with the tensor
// core running asynchronously underneath.  Same math, A/B layouts and MMA order per accumulator as k_conv_s1p / k_conv_tc.
// reference semantics: res_block conv_1 (+nin) and conv_2 (+gate, +shortcut), flow_architecture.py:129-165
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"
#include "tma.cuh"

namespace fn {
using namespace tma;
namespace wg {
static int g_mma_lock = 40;      // 0: off; else the back-off (ns) of a warp waiting for its CTA's MMA-burst lock

constexpr int HDR = 2048;     // [0,512) mbarriers, [512] TMEM base, [1024, 2048) bias (<= 256 floats)

struct Params {
  CUtensorMap tm_x, tm_skip;
  const __nv_bfloat16* wpk;
  const float* bias;
  __nv_bfloat16* y;
  const __nv_bfloat16* res;
  int* status;
  int B, H, W;
  int tiles_u, tiles_v, ntiles;
  int Cout;
  int res_mode;                        // 0 none, 1 plain tile (Cres == F), 2 generic (pool / channel pad), 3 one channel
  int Cres, RH, RW, res_pool;
  int res_u8;                          // res_mode 3: the shortcut is the uint8 boundary image itself
  int mma_lock;                        // 1: serialise the MMA bursts of a CTA's issuing warps
  int s_u, s_v, s_b;                   // the tile stride (grid * NWG) as a mixed-radix number over (tiles_u, tiles_v, B)
};

template <int CLOG, int J, bool GATE, bool SKIP, int NWG>
struct Cfg {
  static constexpr int CH = 1 << CLOG;           // 8-channel chunks of the raw input
  static constexpr int C = 8 * CH;
  static constexpr int F = C;                    // channels written per pixel
  static constexpr int NREAL = GATE ? 2 * C : C;
  static constexpr int N = NREAL < 16 ? 16 : NREAL;
  static constexpr int KS = CH;                  // k-steps per tap (Keff = 2C)
  static constexpr int KSTEPS = 9 * KS + (SKIP ? KS : 0);
  static constexpr int PU = 8 * J + 2, PV = 18, P = PU * PV;      // halo tile of the conv input
  static constexpr int PS = 128 * J;                               // skip operand: centre pixels only (1x1)
  // plane stride (pixels) that keeps the 16-byte stores of 8 neighbouring lanes on distinct banks
  static constexpr int padded(int p) { return CH == 1 ? p : (p + ((CH >= 8 ? 1 : (CH == 4 ? 2 : 4)) - (p & 7) + 8) % 8); }
  static constexpr int PPAD = padded(P), PSPAD = padded(PS);
  static constexpr int r128(int b) { return (b + 127) / 128 * 128; }
  static constexpr int W_BYTES = KSTEPS * N * 32;
  static constexpr int STAGE_BYTES = r128(P * C * 2);
  static constexpr int PLANE_BYTES = r128(2 * CH * PPAD * 16);
  static constexpr int STAGE_S_BYTES = SKIP ? r128(PS * C * 2) : 0;
  static constexpr int PLANE_S_BYTES = SKIP ? r128(2 * CH * PSPAD * 16) : 0;
  static constexpr int WG_BYTES = STAGE_BYTES + PLANE_BYTES + STAGE_S_BYTES + PLANE_S_BYTES;
  static constexpr int OFF_W = HDR;
  static constexpr int OFF_WG = OFF_W + r128(W_BYTES);
  static constexpr int SMEM = OFF_WG + NWG * WG_BYTES;
  static constexpr int ACC_COLS = J * N;                           // per warpgroup
  static constexpr int ALL_COLS = NWG * ACC_COLS;
  static constexpr int TMEM_COLS = ALL_COLS <= 32 ? 32 : ALL_COLS <= 64 ? 64 : ALL_COLS <= 128 ? 128 : ALL_COLS <= 256 ? 256 : 512;
  static constexpr int IN_BYTES = P * C * 2 + (SKIP ? PS * C * 2 : 0);   // bytes one tile's tensor copies deliver
  static_assert(ALL_COLS <= 512, "accumulators exceed TMEM");
  static_assert(N <= 256 && (1 + NWG * (1 + J)) * 8 <= 512, "header layout");
};

__device__ __forceinline__ void wg_barrier(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// stage [pp][C] raw bf16 (TMA box order) -> planes [2C/8][PAD][8] = [elu(x) | elu(-x)]; 128 threads
template <int ITEMS, int CH, int PAD>
__device__ __forceinline__ void transform(uint32_t stage, uint32_t planes, int t128) {
  constexpr int SH = CH == 1 ? 0 : (CH == 2 ? 1 : (CH == 4 ? 2 : 3));
  // rounds of 128 items; the ragged last round starts at thread 64 so that it lands on the warps that issue no MMAs
  constexpr int FULL = ITEMS / 128 * 128;
#pragma unroll 3
  for (int it = t128; it < ITEMS + 64; it += 128) {
    const int idx = it < FULL ? it : FULL + ((it - FULL + 64) & 127);
    if (idx >= ITEMS) continue;
    const int ch = idx & (CH - 1);
    const int pp = idx >> SH;
    const uint4 in = lds128(stage + (uint32_t)idx * 16u);
    uint4 o0, o1;
    concat_elu8(in, o0, o1);
    const uint32_t d0 = planes + (uint32_t)(ch * PAD + pp) * 16u;
    sts128(d0, o0);
    sts128(d0 + (uint32_t)(CH * PAD * 16), o1);
  }
}

__device__ __forceinline__ void lds_f8(uint32_t addr, float (&f)[8]) {
  const uint4 a = lds128(addr), b = lds128(addr + 16u);
  f[0] = __uint_as_float(a.x); f[1] = __uint_as_float(a.y); f[2] = __uint_as_float(a.z); f[3] = __uint_as_float(a.w);
  f[4] = __uint_as_float(b.x); f[5] = __uint_as_float(b.y); f[6] = __uint_as_float(b.z); f[7] = __uint_as_float(b.w);
}


template <int CLOG, int J, bool GATE, bool SKIP, int NWG, int MINB, int RESM>
__global__ void __launch_bounds__(NWG * 128, MINB) k_conv_wg(const __grid_constant__ Params p) {
  using CF = Cfg<CLOG, J, GATE, SKIP, NWG>;
  // no static shared memory in this kernel: the dynamic window starts on a 1024-byte boundary, so the base is a
  // compile-time constant of the shared state space (TMA destinations want 128-byte alignment)
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = ptx::smem_u32(smem);
  auto bar = [&](int i) { return sbase + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + 512);
  float* sbias = reinterpret_cast<float*>(smem + 1024);    // [N]; gate: the v half is pre-multiplied by 0.5 (read back with ld.shared)

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int g = warp >> 2, q = warp & 3;          // warpgroup, TMEM lane quadrant (== warp % 4)
  const int t128 = threadIdx.x & 127;

  if (threadIdx.x == 64) {       // descriptor fetch off the first tile's critical path
    prefetch_tensormap(&p.tm_x);
    if (SKIP) prefetch_tensormap(&p.tm_skip);
  }
  // ---- one-time setup
  if (warp == 0) {
    ptx::tmem_alloc(sbase + 512, (uint32_t)CF::TMEM_COLS);
    ptx::tmem_relinquish();
  } else if (warp == 1 && lane == 0) {
    *reinterpret_cast<volatile int*>(smem + 520) = 0;       // MMA-burst lock (see the issue step)
    ptx::mbar_init(bar(0), 1);
    for (int i = 1; i < 1 + NWG * (1 + J); ++i) ptx::mbar_init(bar(i), 1);
    ptx::fence_mbar_init();
  }
  pdl_wait();                  // everything above overlapped the previous kernel's tail; global memory from here on
  pdl_launch_dependents();
  for (int i = threadIdx.x; i < CF::N; i += NWG * 128) {
    float bv = i < p.Cout ? __ldg(p.bias + i) : 0.f;
    if (GATE && i >= CF::F) bv *= 0.5f;
    sbias[i] = bv;
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(bar(0), (uint32_t)CF::W_BYTES);
    ptx::bulk_g2s(sbase + CF::OFF_W, p.wpk, (uint32_t)CF::W_BYTES, bar(0));
  }

  // ---- this warpgroup's resources
  const uint32_t stage = sbase + CF::OFF_WG + (uint32_t)(g * CF::WG_BYTES);
  const uint32_t planes = stage + CF::STAGE_BYTES;
  const uint32_t stage_s = planes + CF::PLANE_BYTES;
  const uint32_t planes_s = stage_s + CF::STAGE_S_BYTES;
  const uint32_t bar_in = bar(1 + g * (1 + J)), bar_mma = bar_in + 8u;
  const uint32_t acc = tmem_base + (uint32_t)(g * CF::ACC_COLS);

  // tile t -> (tu, tv, b); the loop advances by the constant stride with carries instead of dividing every time
  const int stride = (int)gridDim.x * NWG;
  int t = (int)blockIdx.x * NWG + g;
  int tu = t % p.tiles_u, tv = (t / p.tiles_u) % p.tiles_v, tb = t / (p.tiles_u * p.tiles_v);
  auto advance = [&](int& au, int& av, int& ab) {
    au += p.s_u;
    if (au >= p.tiles_u) { au -= p.tiles_u; ++av; }
    av += p.s_v;
    if (av >= p.tiles_v) { av -= p.tiles_v; ++ab; }
    ab += p.s_b;
  };
  auto issue_loads = [&](int au, int av, int ab) {
    ptx::mbar_arrive_expect_tx(bar_in, (uint32_t)CF::IN_BYTES);
    tma_load_4d(stage, &p.tm_x, bar_in, 0, au * 8 * J - 1, av * 16 - 1, ab);
    if (SKIP) tma_load_4d(stage_s, &p.tm_skip, bar_in, 0, au * 8 * J, av * 16, ab);
  };
  const uint32_t leader = (q < J) ? ptx::elect_one() : 0u;      // one lane of each MMA-issuing warp
  if (q == 0 && leader && t < p.ntiles) issue_loads(tu, tv, tb);
  int nu = tu, nv = tv, nb = tb;           // the tile after this one
  advance(nu, nv, nb);

  const int r = q * 32 + lane;             // accumulator row == TMEM lane
  const int v = r >> 3, u = r & 7;
  const uint32_t lane_acc = acc + ((uint32_t)(q * 32) << 16);
  const uint32_t sb_u = sbase + 1024u, sb_v = sb_u + 4u * (uint32_t)CF::F;
  constexpr uint32_t idesc = ptx::idesc_bf16_f32(128, CF::N);
  bool ok = true, w_ready = false;
  uint32_t par = 0;

  for (; t < p.ntiles; t += stride, par ^= 1u) {
    const int b = tb, y0 = tv * 16, x0 = tu * 8 * J;
    const int oy = y0 + v;
    // ---- shortcut operand: requested first, consumed last (in flight under the transform and the MMAs)
    uint4 rr[RESM == 1 ? J : 1][RESM == 1 ? CF::F / 8 : 1];
    float r1[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int ox = x0 + 8 * j + u;
      const bool inside = oy < p.H && ox < p.W;
      const int pix = (b * p.H + oy) * p.W + ox;
      r1[j] = 0.f;
      if (RESM == 1) {
#pragma unroll
        for (int c8 = 0; c8 < CF::F / 8; ++c8) rr[j][c8] = make_uint4(0u, 0u, 0u, 0u);
        if (inside) {
#pragma unroll
          for (int c8 = 0; c8 < CF::F / 8; ++c8) rr[j][c8] = __ldg(reinterpret_cast<const uint4*>(p.res + (size_t)pix * CF::F) + c8);
        }
      } else if (RESM == 3) {
        // 1-channel shortcut (the boundary image under the first block): lands in the LAST channel (front zero pad, :158-163)
        if (inside) r1[j] = p.res_u8 ? (float)reinterpret_cast<const unsigned char*>(p.res)[pix] : __bfloat162float(p.res[pix]);
      }
    }
    // ---- raw tile(s) -> concat_elu A planes
    ok = ok && ptx::mbar_wait(bar_in, par);
    if (!ok) atomicCAS(p.status, 0, 41);
    transform<CF::P * CF::CH, CF::CH, CF::PPAD>(stage, planes, t128);
    if (SKIP) transform<CF::PS * CF::CH, CF::CH, CF::PSPAD>(stage_s, planes_s, t128);
    ptx::fence_proxy_async_smem();          // this thread's plane writes -> tensor-core proxy
    ptx::tc_fence_before_sync();            // this thread's TMEM reads of the previous tile are done
    wg_barrier(g);
    // ---- warp j of the warpgroup issues sub-tile j's MMAs (independent accumulators, so the J programs run side by
    // side); the first warp also starts the next tile's tensor copies into the now free stage
    if (q < J) {
      ptx::tc_fence_after_sync();
      if (q == 0 && leader && t + stride < p.ntiles) issue_loads(nu, nv, nb);
      if (!w_ready) {
        ok = ok && ptx::mbar_wait(bar(0), 0);
        if (!ok) atomicCAS(p.status, 0, 42);
        w_ready = true;
      }
      __syncwarp();
      // the warp runs the unrolled program convergently (descriptor arithmetic on the uniform datapath, immediate offsets)
      const uint64_t a_tmpl = ptx::smem_desc(planes + (uint32_t)(q * 8 * 16), CF::PPAD * 16, CF::PU * 16);
      const uint64_t s_tmpl = ptx::smem_desc(planes_s + (uint32_t)(q * 8 * 16), CF::PSPAD * 16, 8 * J * 16);
      const uint64_t b_tmpl = ptx::smem_desc(sbase + CF::OFF_W, CF::N * 16, 128);
      const uint32_t d_tmem = acc + (uint32_t)(q * CF::N);
      if (leader) {
        // one sub-tile's MMAs as a burst: with every issuing warp of the SM feeding the MMA unit round-robin, each
        // program of nine 39-clock instructions took 3.5 k clocks to get through (processor sharing); serialised per CTA
        // (two CTAs per SM keep the unit fed) a program completes in ~0.5 k once it starts, and its epilogue with it
        if (p.mma_lock) {
          int* lock = reinterpret_cast<int*>(smem + 520);
          while (atomicCAS(lock, 0, 1) != 0) __nanosleep((unsigned)p.mma_lock);
        }
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
          for (int kk = 0; kk < CF::KS; ++kk) {
            const uint32_t a_off = (uint32_t)(((tap / 3) * CF::PU + (tap % 3)) * 16 + kk * 2 * CF::PPAD * 16) >> 4;
            const uint32_t b_off = (uint32_t)((tap * CF::KS + kk) * CF::N * 32) >> 4;
            ptx::mma_bf16_ss(d_tmem, a_tmpl + a_off, b_tmpl + b_off, idesc, (tap | kk) ? 1u : 0u);
          }
        }
        if (SKIP) {
#pragma unroll
          for (int kk = 0; kk < CF::KS; ++kk) {
            const uint32_t a_off = (uint32_t)(kk * 2 * CF::PSPAD * 16) >> 4;
            const uint32_t b_off = (uint32_t)((9 * CF::KS + kk) * CF::N * 32) >> 4;
            ptx::mma_bf16_ss(d_tmem, s_tmpl + a_off, b_tmpl + b_off, idesc, 1u);
          }
        }
        ptx::mma_commit(bar_mma + 8u * (uint32_t)q);
        if (p.mma_lock) atomicExch(reinterpret_cast<int*>(smem + 520), 0);
      }
      __syncwarp();
    }
    // ---- epilogue: TMEM -> bias, gate * sigmoid, shortcut -> bf16 NHWC (16 bytes per lane and 8 channels)
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int ox = x0 + 8 * j + u;
      const int pix = (b * p.H + oy) * p.W + ox;
      ok = ok && ptx::mbar_wait(bar_mma + 8u * (uint32_t)j, par);
      if (!ok) atomicCAS(p.status, 0, 43);
      ptx::tc_fence_after_sync();
      if (RESM == 1) {      // keep the compiler from unpacking (= waiting for) the shortcut loads before this point
#pragma unroll
        for (int c8 = 0; c8 < CF::F / 8; ++c8)
          asm volatile("" : "+r"(rr[j][c8].x), "+r"(rr[j][c8].y), "+r"(rr[j][c8].z), "+r"(rr[j][c8].w));
      }
      const bool live = ok && oy < p.H && ox < p.W;
      const uint32_t col0 = lane_acc + (uint32_t)(j * CF::N);
      __nv_bfloat16* ydst = p.y + (size_t)pix * CF::F;
#pragma unroll
      for (int c8 = 0; c8 < CF::F / 8; ++c8) {
        const int c0 = 8 * c8;
        float o[8], bu[8];
        lds_f8(sb_u + 4u * (uint32_t)c0, bu);
        if (GATE) {
          float a[8], gt[8], bv[8];
          lds_f8(sb_v + 4u * (uint32_t)c0, bv);
          ptx::tmem_ld8x2(col0 + c0, col0 + CF::F + c0, a, gt);
          // u * sigmoid(v) with sigmoid(z) = 0.5 + 0.5 tanh(z/2); the v bias is stored pre-halved
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float th = tanh_approx(fmaf(gt[i], 0.5f, bv[i]));
            o[i] = (a[i] + bu[i]) * fmaf(th, 0.5f, 0.5f);
          }
          if (RESM == 1) {
            bf16x8 rv;
            rv.v = rr[j][c8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] += __bfloat162float(rv.h[i]);
          } else if (RESM == 2) {
            if (live) shortcut_add8(p.res, p.Cres, p.RH, p.RW, p.res_pool, CF::F, b, oy, ox, c0, o);
          } else if (RESM == 3) {
            if (c0 == CF::F - 8) o[7] += r1[j];
          }
        } else {
          ptx::tmem_ld8(col0 + c0, o);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] += bu[i];
        }
        if (live) {
          bf16x8 st;
#pragma unroll
          for (int i = 0; i < 8; ++i) st.h[i] = __float2bfloat16(o[i]);
          *reinterpret_cast<uint4*>(ydst + c0) = st.v;
        }
      }
    }
    tu = nu; tv = nv; tb = nb;
    advance(nu, nv, nb);
  }

  // ---- teardown
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, (uint32_t)CF::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------ host side

static int g_enable = -1;     // -1: read FLOWNET_WG on first use (default 1); 0 off; 1 where measured faster than k_conv_s1p; 2 every shape it has
static int g_j16 = -1;        // sub-tiles per warpgroup tile on the 16-channel level (FLOWNET_WG_J16, default 1)

template <int CLOG, int J, bool GATE, bool SKIP, int NWG, int MINB, int RESM>
static int launch_m(const fn_conv_desc& d, int* status, cudaStream_t stream) {
  using CF = Cfg<CLOG, J, GATE, SKIP, NWG>;
  static_assert(CF::SMEM * MINB + 1024 * MINB <= 227 * 1024, "shared memory of the resident CTAs exceeds the SM");
  static_assert(CF::TMEM_COLS * MINB <= 512, "TMEM of the resident CTAs exceeds the SM");
  Params p;
  memset(&p, 0, sizeof(p));
  if (!make_map_nhwc(&p.tm_x, d.x, d.B, d.H, d.W, CF::C, CF::PU, CF::PV)) {
    set_error("conv_wg: cuTensorMapEncodeTiled(x) failed");
    return FN_E_UNSUPPORTED;
  }
  if (SKIP && !make_map_nhwc(&p.tm_skip, d.skip, d.B, d.SH, d.SW, CF::C, 8 * J, 16)) {
    set_error("conv_wg: cuTensorMapEncodeTiled(skip) failed");
    return FN_E_UNSUPPORTED;
  }
  p.res_mode = RESM;
  p.res = (const __nv_bfloat16*)d.res;
  p.res_u8 = d.res_u8;
  p.wpk = (const __nv_bfloat16*)d.w_tc;
  p.bias = d.bias;
  p.y = (__nv_bfloat16*)d.y;
  p.status = status;
  p.B = d.B; p.H = d.H; p.W = d.W;
  p.tiles_u = (d.W + 8 * J - 1) / (8 * J);
  p.tiles_v = (d.H + 15) / 16;
  p.ntiles = d.B * p.tiles_u * p.tiles_v;
  p.Cout = d.Cout;
  p.Cres = d.Cres; p.RH = d.RH; p.RW = d.RW; p.res_pool = d.res_pool;

  auto kern = k_conv_wg<CLOG, J, GATE, SKIP, NWG, MINB, RESM>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(conv_wg): %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  int grid = (p.ntiles + NWG - 1) / NWG;
  if (grid > 148 * MINB) grid = 148 * MINB;
  const int stride = grid * NWG;
  p.s_u = stride % p.tiles_u;
  p.s_v = (stride / p.tiles_u) % p.tiles_v;
  p.s_b = stride / (p.tiles_u * p.tiles_v);
  // the burst lock pays under contention (batch 64: -3 ... -13 % per launch); with at most a tile or two per warpgroup (batch 8)
  // its two atomics per burst only add latency (+3 % on the whole forward)
  p.mma_lock = p.ntiles >= 3 * stride ? g_mma_lock : 0;
  launch_pdl(kern, dim3(grid), dim3(NWG * 128), (size_t)CF::SMEM, stream, p);
  count_launch();
  return check_launch("k_conv_wg");
}

// shortcut operand: 0 none, 1 plain tile (Cres == F), 2 generic (pool / channel pad), 3 one channel
template <int CLOG, int J, bool GATE, bool SKIP, int NWG, int MINB>
static int launch(const fn_conv_desc& d, int* status, cudaStream_t stream) {
  if (GATE && d.res != nullptr) {
    if (!d.res_pool && d.Cres == 8 << CLOG && d.RH == d.H && d.RW == d.W) return launch_m<CLOG, J, GATE, SKIP, NWG, MINB, GATE ? 1 : 0>(d, status, stream);
    if (!d.res_pool && d.Cres == 1 && d.RH == d.H && d.RW == d.W) return launch_m<CLOG, J, GATE, SKIP, NWG, MINB, GATE ? 3 : 0>(d, status, stream);
    return launch_m<CLOG, J, GATE, SKIP, NWG, MINB, GATE ? 2 : 0>(d, status, stream);
  }
  return launch_m<CLOG, J, GATE, SKIP, NWG, MINB, 0>(d, status, stream);
}

}  // namespace wg

int* tc_status_word();   // conv_tc.cu: lazily allocated device status word

// Returns true if a warpgroup kernel exists for this launch (and runs it unless dry_run: rc = its return code).
bool conv_wg_try(const fn_conv_desc& d, cudaStream_t stream, bool dry_run, int& rc) {
  using namespace wg;
  rc = 0;
  if (g_enable < 0) {
    const char* e = getenv("FLOWNET_WG");
    g_enable = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1;
  }
  if (g_j16 < 0) {
    const char* e = getenv("FLOWNET_WG_J16");
    g_j16 = (e && e[0] == '2') ? 2 : 1;
  }
  if (!g_enable) return false;
  if (d.ksize != 3 || d.stride != 1 || d.prologue != FN_PRO_CONCAT_ELU || d.w_tc == nullptr) return false;
  if (d.keep_prob < 1.0f) return false;                       // training prologue: k_conv_s1p
  const bool gate = d.epilogue == FN_EPI_GATE_RES;
  if (!gate && d.epilogue != FN_EPI_BIAS) return false;       // the gate adjoint stays on k_conv_s1p
  if (d.Cout != (gate ? 2 * d.Cin : d.Cin)) return false;
  const bool skip = d.skip != nullptr;
  if (skip && (gate || d.Cskip != d.Cin)) return false;
  if (d.H < 16 || d.W < 8) return false;
  if (d.res_u8 && !(d.res && d.Cres == 1 && !d.res_pool && d.RH == d.H && d.RW == d.W)) return false;
  if (get_encode() == nullptr) return false;
  const int clog = d.Cin == 8 ? 0 : d.Cin == 16 ? 1 : d.Cin == 32 ? 2 : -1;
  if (clog < 0) return false;
  if (clog == 2 && gate) return false;                        // 147 KB of weights: k_conv_s1p (one CTA per SM)
  // A/B at batch 64 (tools/ab_wg.py, profiles/): the plain 16-channel conv_1 and the 32-channel level stay on k_conv_s1p
  if (g_enable == 1 && ((clog == 1 && !gate && !skip) || clog == 2)) return false;
  const int j = clog == 0 ? 2 : clog == 1 ? g_j16 : 1;
#define WG_CASE(CL, JJ, G, S, NW, MB)                                                   \
  if (clog == CL && j == JJ && gate == G && skip == S) {                                 \
    if (!dry_run) {                                                                      \
      int* st = tc_status_word();                                                        \
      rc = !st ? (int)cudaErrorMemoryAllocation : launch<CL, JJ, G, S, NW, MB>(d, st, stream); \
    }                                                                                    \
    return true;                                                                         \
  }
  // warpgroups per CTA x CTAs per SM, measured at batch 64 (round 2, with the MMA-burst lock): the 8-channel level wants more
  // resident warpgroups than registers comfortably allow -- gate 4x2 42.4 us, 3x3 38.5 (53 registers, 32 bytes of spill),
  // 3x4 46.6, 2x5 51.8, 5x2 46.3; plain 4x2 38.4, 3x3 35.7, 3x4 33.4; with the skip operand 3x2 43.0, 2x3 41.3.  The
  // 16-channel level stays at 4x2 (gate 22.7 us against 24.0 for 3x3).
  WG_CASE(0, 2, true, false, 3, 3) WG_CASE(0, 2, false, false, 3, 4) WG_CASE(0, 2, false, true, 2, 3)
  WG_CASE(1, 1, true, false, 4, 2) WG_CASE(1, 1, false, false, 4, 2) WG_CASE(1, 1, false, true, 3, 2)
  WG_CASE(1, 2, true, false, 4, 1) WG_CASE(1, 2, false, false, 4, 1) WG_CASE(1, 2, false, true, 3, 1)
  WG_CASE(2, 1, false, false, 4, 1) WG_CASE(2, 1, false, true, 3, 1)
#undef WG_CASE
  return false;
}

}  // namespace fn

extern "C" {
// test / timing hook: 0 sends the stride-1 concat_elu convolutions back to k_conv_s1p, 1 (default) uses the warpgroup kernel
// where it measured faster, 2 wherever it has an instantiation
void fn_debug_set_wg(int v) { fn::wg::g_enable = v < 0 ? 0 : v > 2 ? 2 : v; }
void fn_debug_set_wg_j16(int v) { fn::wg::g_j16 = v == 2 ? 2 : 1; }
void fn_debug_set_wg_lock(int v) { fn::wg::g_mma_lock = v < 0 ? 0 : v; }
}
