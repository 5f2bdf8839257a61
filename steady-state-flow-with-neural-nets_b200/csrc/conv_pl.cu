// Stride-1 3x3 convolutions of flow_net on MMA-READY activations ("planes"), inference path, C in {8,16,32}
// (the 128x256, 64x128 and 32x64 levels; reference flow_architecture.py:129-165: conv_1 (+ nin skip) and conv_2 + gate +
// shortcut of a res_block).
//
// Round 1/2 measurements (ncu of k_conv_wg): with the concat_elu prologue computed by the CONSUMER,
// every tile goes TMA -> CUDA-core transform -> MMA -> epilogue, the shared-memory pipe is ~70 % busy (9 tap re-reads by
// the tensor core + stage read + two plane writes by the transform) and the stage latencies add.  Here the PRODUCING
// kernel's epilogue stores [elu(y) | elu(-y)] once, in the layout the UMMA A operand wants:
//
//   planes tensor   bf16 [B][2C/8][H][W][8]      plane q < C/8: elu(t[..., 8q:8q+8]),  plane C/8 + q: elu(-t[..., 8q:8q+8])
//
// so one tensor copy (4-D map over 4-byte elements, box = (8J+2) x 18 pixels x 2C/8 planes, zero fill outside the image
// = TF 'SAME', and concat_elu(0) = 0) lands a halo tile in shared memory exactly as the K-major SWIZZLE_NONE operand
// [plane][row][pixel][8 ch]: a filter tap is a descriptor start offset, as in every other kernel of this library.
// The kernel is a plain three-role pipeline with nothing between the copy and the MMA:
//
//   warp NE     producer   tensor copies of tile i+NST-1.. into an NST-deep ring (full / empty mbarriers)
//   warp NE+1   issuer     one elected lane: unrolled tcgen05.mma program of a tile (J sub-tiles of 128 pixels x 9 taps
//                          x C/8 k-steps, + the nin skip's centre tap), accumulators NACC deep in TMEM; commit -> ring
//                          slot free, commit -> accumulator full
//   warps 0..NE-1 epilogue TMEM -> bias / gate * sigmoid / shortcut -> bf16, stored as raw NHWC (for shortcuts, resampling
//                          convs, the tail) and / or as the planes of the NEXT conv
//
// Same weights image (fn_pack_conv_weights_tc), MMA order per accumulator and rounding points as k_conv_wg / k_conv_s1p:
// the raw output is bit-identical to theirs, and the planes are concat_elu of that bf16 output, as their transform computes.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"
#include "tma.cuh"

namespace fn {
using namespace tma;
namespace pl {

constexpr int HDR = 2048;     // [0,512) mbarriers, [512] TMEM base, [1024, 2048) bias (<= 256 floats)

struct Params {
  CUtensorMap tm_x, tm_skip;
  const __nv_bfloat16* wpk;
  const float* bias;
  __nv_bfloat16* y;                    // raw NHWC output or null
  __nv_bfloat16* yp;                   // planes output or null
  const __nv_bfloat16* res;
  int* status;
  int B, H, W;
  int tiles_u, tiles_v, ntiles;
  int Cout;
  int res_mode;                        // 0 none, 1 plain tile (Cres == F), 2 generic (pool / channel pad), 3 one channel
  int Cres, RH, RW, res_pool;
  int knock;                           // debug (tools/ab_planes.py): 1 no MMAs, 2 no stores, 4 no tensor copies
};

template <int CLOG, int J, bool GATE, bool SKIP, int NST, int NACC, int NE, int NI>
struct Cfg {
  static constexpr int CH = 1 << CLOG;           // 8-channel chunks of the raw tensor
  static constexpr int C = 8 * CH;
  static constexpr int F = C;                    // channels written per pixel
  static constexpr int NREAL = GATE ? 2 * C : C;
  static constexpr int N = NREAL < 16 ? 16 : NREAL;
  static constexpr int KS = CH;                  // k-steps per tap (K = 2C)
  static constexpr int KSTEPS = 9 * KS + (SKIP ? KS : 0);
  static constexpr int PU = 8 * J + 2, PV = 18, P = PU * PV;      // halo tile of the conv input
  static constexpr int SU = 8 * J, SV = 16, PS = SU * SV;         // skip operand: centre pixels only (1x1)
  static constexpr int r128(int b) { return (b + 127) / 128 * 128; }
  static constexpr int W_BYTES = KSTEPS * N * 32;
  static constexpr int X_BYTES = 2 * CH * P * 16;
  static constexpr int S_BYTES = SKIP ? 2 * CH * PS * 16 : 0;
  static constexpr int IN_BYTES = X_BYTES + S_BYTES;
  static constexpr int OFF_S = r128(X_BYTES);                      // skip tile inside a ring slot
  static constexpr int STAGE_BYTES = OFF_S + r128(S_BYTES);
  static constexpr int OFF_W = HDR;
  static constexpr int OFF_ST = OFF_W + r128(W_BYTES);
  static constexpr int SMEM = OFF_ST + NST * STAGE_BYTES;
  static constexpr int ACC_COLS = J * N;
  static constexpr int ALL_COLS = NACC * ACC_COLS;
  static constexpr int TMEM_COLS = ALL_COLS <= 32 ? 32 : ALL_COLS <= 64 ? 64 : ALL_COLS <= 128 ? 128 : ALL_COLS <= 256 ? 256 : 512;
  static constexpr int EG = NE / 4;                                // epilogue warps per TMEM lane quadrant
  static constexpr int WARPS = NE + 1 + NI;                        // epilogue, producer, issuers
  static constexpr int ITEMS = J * (F / 8);                        // (sub-tile, 8-channel group) work items of a tile
  static constexpr int CHUNK = ITEMS < 4 ? ITEMS : 4;              // items whose TMEM loads are in flight together
  static_assert(ITEMS % CHUNK == 0, "items per tile must be a multiple of the chunk");
  // every mbarrier has ONE waiter that takes its phases in order (a parity wait cannot tell phase k from phase k+2):
  // accumulator a is always drained by epilogue group a % EG and filled by issuer a % NI, ring slot s read by issuer s % NI
  static_assert(NACC % EG == 0 && NACC % NI == 0 && NST % NI == 0, "barrier ownership");
  static constexpr int BW = 0, FULL = 1, EMPTY = FULL + NST, ACCF = EMPTY + NST, ACCE = ACCF + NACC, NBAR = ACCE + NACC;
  static_assert(ALL_COLS <= 512, "accumulators exceed TMEM");
  static_assert(N <= 256 && NBAR * 8 <= 512, "header layout");
  static_assert(NE % 4 == 0 && SMEM <= 227 * 1024, "configuration");
  static_assert((X_BYTES % 128) == 0, "skip tile must start on a 128-byte boundary");
};

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void lds_f8(uint32_t addr, float (&f)[8]) {
  const uint4 a = lds128(addr), b = lds128(addr + 16u);
  f[0] = __uint_as_float(a.x); f[1] = __uint_as_float(a.y); f[2] = __uint_as_float(a.z); f[3] = __uint_as_float(a.w);
  f[4] = __uint_as_float(b.x); f[5] = __uint_as_float(b.y); f[6] = __uint_as_float(b.z); f[7] = __uint_as_float(b.w);
}


template <int CLOG, int J, bool GATE, bool SKIP, int NST, int NACC, int NE, int NI>
__global__ void __launch_bounds__((NE + 1 + NI) * 32, 1) k_conv_pl(const __grid_constant__ Params p) {
  using CF = Cfg<CLOG, J, GATE, SKIP, NST, NACC, NE, NI>;
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = ptx::smem_u32(smem);
  auto bar = [&](int i) { return sbase + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + 512);
  float* sbias = reinterpret_cast<float*>(smem + 1024);    // [N]; gate: the v half is pre-multiplied by 0.5

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  // ---- one-time setup
  if (warp == NE + 1) {
    ptx::tmem_alloc(sbase + 512, (uint32_t)CF::TMEM_COLS);
    ptx::tmem_relinquish();
  } else if (warp == NE && lane == 0) {
    ptx::mbar_init(bar(CF::BW), 1);
    for (int i = 0; i < NST; ++i) {
      ptx::mbar_init(bar(CF::FULL + i), 1);
      ptx::mbar_init(bar(CF::EMPTY + i), 1);
    }
    for (int i = 0; i < NACC; ++i) {
      ptx::mbar_init(bar(CF::ACCF + i), 1);
      ptx::mbar_init(bar(CF::ACCE + i), 4);       // the four warps of the group that drained it
    }
    ptx::fence_mbar_init();
  }
  pdl_wait();                  // everything above overlapped the previous kernel's tail; global memory from here on
  pdl_launch_dependents();
  for (int i = threadIdx.x; i < CF::N; i += CF::WARPS * 32) {
    float bv = i < p.Cout ? __ldg(p.bias + i) : 0.f;
    if (GATE && i >= CF::F) bv *= 0.5f;
    sbias[i] = bv;
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_img = p.tiles_u * p.tiles_v;
  bool ok = true;

  if (warp == NE) {
    // =========================== producer ===========================
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(bar(CF::BW), (uint32_t)CF::W_BYTES);
      for (int off = 0; off < CF::W_BYTES; off += 32768) {
        const int n = CF::W_BYTES - off < 32768 ? CF::W_BYTES - off : 32768;
        ptx::bulk_g2s(sbase + CF::OFF_W + (uint32_t)off, reinterpret_cast<const unsigned char*>(p.wpk) + off, (uint32_t)n, bar(CF::BW));
      }
      int i = 0;
      for (int t = (int)blockIdx.x; t < p.ntiles; t += (int)gridDim.x, ++i) {
        const int s = i % NST;
        const uint32_t ph = (uint32_t)(i / NST) & 1u;
        ok = ok && ptx::mbar_wait(bar(CF::EMPTY + s), ph ^ 1u);
        if (!ok) { atomicCAS(p.status, 0, 61); break; }
        const int b = t / tiles_img, rem = t - b * tiles_img;
        const int tv = rem / p.tiles_u, tu = rem - tv * p.tiles_u;
        const uint32_t dst = sbase + CF::OFF_ST + (uint32_t)(s * CF::STAGE_BYTES);
        if (p.knock & 4) { ptx::mbar_arrive(bar(CF::FULL + s)); continue; }
        ptx::mbar_arrive_expect_tx(bar(CF::FULL + s), (uint32_t)CF::IN_BYTES);
        tma_load_4d(dst, &p.tm_x, bar(CF::FULL + s), (tu * 8 * J - 1) * 4, tv * 16 - 1, 0, b);
        if (SKIP) tma_load_4d(dst + (uint32_t)CF::OFF_S, &p.tm_skip, bar(CF::FULL + s), tu * 8 * J * 4, tv * 16, 0, b);
      }
    }
    __syncwarp();
  } else if (warp > NE) {
    // =========================== issuers: warp NE+1+w takes tiles w, w+NI, ... ===========================
    // One elected lane runs the unrolled program of a tile.  The J sub-tiles' accumulators are visited in rotation
    // (tap-major), so consecutive MMAs never wait for each other's accumulator; per accumulator the order is still
    // tap 0..8 x k-step, then the skip.  Several issuer warps hide the per-tile barrier round trips of a single thread.
    const int wi = warp - (NE + 1);
    const uint32_t leader = ptx::elect_one();
    constexpr uint32_t idesc = ptx::idesc_bf16_f32(128, CF::N);
    ok = ptx::mbar_wait(bar(CF::BW), 0);
    if (!ok) atomicCAS(p.status, 0, 62);
    const uint64_t b_tmpl = ptx::smem_desc(sbase + CF::OFF_W, CF::N * 16, 128);
    for (int i = wi, t = (int)blockIdx.x + wi * (int)gridDim.x; t < p.ntiles && ok; i += NI, t += NI * (int)gridDim.x) {
      const int s = i % NST, a = i % NACC;
      const uint32_t ph = (uint32_t)(i / NST) & 1u, pa = (uint32_t)(i / NACC) & 1u;
      ok = ok && ptx::mbar_wait(bar(CF::ACCE + a), pa ^ 1u);
      ok = ok && ptx::mbar_wait(bar(CF::FULL + s), ph);
      if (!ok) { atomicCAS(p.status, 0, 63); break; }
      ptx::tc_fence_after_sync();
      if (leader && (p.knock & 1)) {
        ptx::mma_commit(bar(CF::EMPTY + s));
        ptx::mma_commit(bar(CF::ACCF + a));
      } else if (leader) {
        const uint32_t st = sbase + CF::OFF_ST + (uint32_t)(s * CF::STAGE_BYTES);
        const uint32_t acc = tmem_base + (uint32_t)(a * CF::ACC_COLS);
        const uint64_t a_tmpl = ptx::smem_desc(st, CF::P * 16, CF::PU * 16);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
          for (int kk = 0; kk < CF::KS; ++kk) {
#pragma unroll
            for (int j = 0; j < J; ++j) {
              const uint32_t a_off = (uint32_t)(j * 8 * 16 + ((tap / 3) * CF::PU + (tap % 3)) * 16 + kk * 2 * CF::P * 16) >> 4;
              const uint32_t b_off = (uint32_t)((tap * CF::KS + kk) * CF::N * 32) >> 4;
              ptx::mma_bf16_ss(acc + (uint32_t)(j * CF::N), a_tmpl + a_off, b_tmpl + b_off, idesc, (tap | kk) ? 1u : 0u);
            }
          }
        }
        if (SKIP) {
          const uint64_t s_tmpl = ptx::smem_desc(st + (uint32_t)CF::OFF_S, CF::PS * 16, CF::SU * 16);
#pragma unroll
          for (int kk = 0; kk < CF::KS; ++kk) {
#pragma unroll
            for (int j = 0; j < J; ++j) {
              const uint32_t a_off = (uint32_t)(j * 8 * 16 + kk * 2 * CF::PS * 16) >> 4;
              const uint32_t b_off = (uint32_t)((9 * CF::KS + kk) * CF::N * 32) >> 4;
              ptx::mma_bf16_ss(acc + (uint32_t)(j * CF::N), s_tmpl + a_off, b_tmpl + b_off, idesc, 1u);
            }
          }
        }
        ptx::mma_commit(bar(CF::EMPTY + s));       // the ring slot is free once these MMAs have read it
        ptx::mma_commit(bar(CF::ACCF + a));
      }
      __syncwarp();
    }
  } else {
    // =========================== epilogue: NE warps = EG groups of four ===========================
    // Group eg (one warp per TMEM lane quadrant) takes every EG-th tile of this CTA whole: the per-tile overhead (barrier
    // wait, tile coordinates, hand-back) is paid once per ITEMS work items, the items of a chunk are independent chains
    // (TMEM loads issued together, one wait), and EG tiles are in flight per quadrant.
    const int qd = warp & 3, eg = warp >> 2;
    const int r = qd * 32 + lane;             // accumulator row == TMEM lane
    const int v = r >> 3, u = r & 7;
    const uint32_t lane_base = tmem_base + ((uint32_t)(qd * 32) << 16);
    const uint32_t sb_u = sbase + 1024u, sb_v = sb_u + 4u * (uint32_t)CF::F;
    const size_t plane_px = (size_t)p.H * p.W;
    constexpr int F8 = CF::F / 8;
    constexpr int CHUNK = CF::CHUNK, NCHUNK = CF::ITEMS / CHUNK;
    for (int i = eg, t = (int)blockIdx.x + eg * (int)gridDim.x; t < p.ntiles; i += CF::EG, t += CF::EG * (int)gridDim.x) {
      const int a = i % NACC;
      const uint32_t pa = (uint32_t)(i / NACC) & 1u;
      const int b = t / tiles_img, rem = t - b * tiles_img;
      const int tv = rem / p.tiles_u, tu = rem - tv * p.tiles_u;
      const int oy = tv * 16 + v, x0 = tu * 8 * J;
      const bool row_live = oy < p.H;
      const size_t row_pix = ((size_t)b * p.H + oy) * p.W;
#pragma unroll
      for (int ch = 0; ch < NCHUNK; ++ch) {
        // ---- shortcut operand of the chunk: requested before the accumulator is awaited
        uint4 rr[GATE ? CHUNK : 1];
        float r1[GATE ? CHUNK : 1];
        if (GATE) {
#pragma unroll
          for (int k = 0; k < CHUNK; ++k) {
            const int item = ch * CHUNK + k;
            const int j = item / F8, c8 = item % F8;
            const int ox = x0 + 8 * j + u;
            rr[k] = make_uint4(0u, 0u, 0u, 0u);
            r1[k] = 0.f;
            if (row_live && ox < p.W) {
              if (p.res_mode == 1) rr[k] = __ldg(reinterpret_cast<const uint4*>(p.res + (row_pix + ox) * CF::F) + c8);
              else if (p.res_mode == 3) r1[k] = __bfloat162float(p.res[row_pix + ox]);
            }
          }
        }
        if (ch == 0) {
          ok = ok && ptx::mbar_wait(bar(CF::ACCF + a), pa);
          if (!ok) atomicCAS(p.status, 0, 64);
          ptx::tc_fence_after_sync();
        }
        // ---- the chunk's TMEM loads, one wait
        uint32_t acc[CHUNK][GATE ? 16 : 8];
#pragma unroll
        for (int k = 0; k < CHUNK; ++k) {
          const int item = ch * CHUNK + k;
          const int j = item / F8, c8 = item % F8;
          const uint32_t col0 = lane_base + (uint32_t)(a * CF::ACC_COLS + j * CF::N + 8 * c8);
          if constexpr (GATE) ptx::tmem_ld8x2_issue(col0, col0 + CF::F, acc[k]);
          else ptx::tmem_ld8_issue(col0, acc[k]);
        }
        if constexpr (GATE) ptx::tmem_wait_pin16(acc[0]); else ptx::tmem_wait_pin8(acc[0]);
#pragma unroll
        for (int k = 1; k < CHUNK; ++k) {
          if constexpr (GATE) ptx::pin16(acc[k]); else ptx::pin8(acc[k]);
        }
        if (ch == NCHUNK - 1) {      // the accumulator is in registers: hand it back before the arithmetic
          ptx::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(bar(CF::ACCE + a));
        }
#pragma unroll
        for (int k = 0; k < CHUNK; ++k) {
          const int item = ch * CHUNK + k;
          const int j = item / F8, c8 = item % F8;
          const int c0 = 8 * c8;
          const int ox = x0 + 8 * j + u;
          const bool live = ok && row_live && ox < p.W && !(p.knock & 2);
          const size_t pix = row_pix + ox;
          float o[8], bu[8];
          lds_f8(sb_u + 4u * (uint32_t)c0, bu);
          if constexpr (GATE) {
            float bv[8];
            lds_f8(sb_v + 4u * (uint32_t)c0, bv);
            // u * sigmoid(v) with sigmoid(z) = 0.5 + 0.5 tanh(z/2); the v bias is stored pre-halved
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float th = tanh_approx(fmaf(__uint_as_float(acc[k][8 + e]), 0.5f, bv[e]));
              o[e] = (__uint_as_float(acc[k][e]) + bu[e]) * fmaf(th, 0.5f, 0.5f);
            }
            if (p.res_mode == 1) {
              bf16x8 rv;
              rv.v = rr[k];
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] += __bfloat162float(rv.h[e]);
            } else if (p.res_mode == 2) {
              if (live) shortcut_add8(p.res, p.Cres, p.RH, p.RW, p.res_pool, CF::F, b, oy, ox, c0, o);
            } else if (p.res_mode == 3) {
              if (c0 == CF::F - 8) o[7] += r1[k];
            }
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(acc[k][e]) + bu[e];
          }
          if (live) {
            bf16x8 st;
#pragma unroll
            for (int e = 0; e < 8; ++e) st.h[e] = __float2bfloat16(o[e]);
            if (p.y) *reinterpret_cast<uint4*>(p.y + pix * CF::F + c0) = st.v;
            if (p.yp) {
              uint4 e0, e1;
              concat_elu8(st.v, e0, e1);
              __nv_bfloat16* d0 = p.yp + (((size_t)b * (2 * CF::CH) + c8) * plane_px + (size_t)oy * p.W + ox) * 8;
              *reinterpret_cast<uint4*>(d0) = e0;
              *reinterpret_cast<uint4*>(d0 + (size_t)CF::CH * plane_px * 8) = e1;
            }
          }
        }
      }
    }
  }

  // ---- teardown
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == NE + 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, (uint32_t)CF::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------ host side

// planes tensor [B][planes][H][W][8] bf16 viewed as 4-byte elements (a pixel's 8 channels = 4 elements): the box limit of
// 256 elements per dimension then allows rows of up to 64 pixels
static bool make_map_planes(CUtensorMap* tm, const void* base, int B, int H, int W, int planes, int bw, int bh) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)W * 4, (cuuint64_t)H, (cuuint64_t)planes, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)planes * H * W * 16};
  const cuuint32_t box[4] = {(cuuint32_t)bw * 4, (cuuint32_t)bh, (cuuint32_t)planes, 1u};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int g_enable = -1;     // -1: read FLOWNET_PL on first use (default 1); 0: refuse planes operands
static int g_cap = 0;         // debug: cap on the grid size (0 = none)
static int g_ne = 0;          // debug: force the epilogue warps (8 or 16; 0 = the shape's default)
static int g_ni = 0;          // debug: force the issuer warps (1, 2 or 4; 0 = the shape's default)
static int g_knock = 0;       // debug: pipeline stages switched off (Params::knock)
static int g_j = 0;           // debug: force the tile width (0 = first instantiation of the shape)

template <int CLOG, int J, bool GATE, bool SKIP, int NST, int NACC, int NE, int NI>
static int launch(const fn_conv_desc& d, int* status, cudaStream_t stream) {
  using CF = Cfg<CLOG, J, GATE, SKIP, NST, NACC, NE, NI>;
  Params p;
  memset(&p, 0, sizeof(p));
  if (!make_map_planes(&p.tm_x, d.x, d.B, d.H, d.W, 2 * CF::CH, CF::PU, CF::PV)) {
    set_error("conv_pl: cuTensorMapEncodeTiled(x planes) failed");
    return FN_E_UNSUPPORTED;
  }
  if (SKIP && !make_map_planes(&p.tm_skip, d.skip, d.B, d.SH, d.SW, 2 * CF::CH, CF::SU, CF::SV)) {
    set_error("conv_pl: cuTensorMapEncodeTiled(skip planes) failed");
    return FN_E_UNSUPPORTED;
  }
  p.res = (const __nv_bfloat16*)d.res;
  p.res_mode = 0;
  if (GATE && d.res != nullptr) {
    if (!d.res_pool && d.Cres == CF::F && d.RH == d.H && d.RW == d.W) p.res_mode = 1;
    else if (!d.res_pool && d.Cres == 1 && d.RH == d.H && d.RW == d.W) p.res_mode = 3;
    else p.res_mode = 2;
  }
  p.wpk = (const __nv_bfloat16*)d.w_tc;
  p.bias = d.bias;
  p.y = (__nv_bfloat16*)d.y;
  p.yp = (__nv_bfloat16*)d.y_planes;
  p.status = status;
  p.B = d.B; p.H = d.H; p.W = d.W;
  p.tiles_u = (d.W + 8 * J - 1) / (8 * J);
  p.tiles_v = (d.H + 15) / 16;
  p.ntiles = d.B * p.tiles_u * p.tiles_v;
  p.Cout = d.Cout;
  p.Cres = d.Cres; p.RH = d.RH; p.RW = d.RW; p.res_pool = d.res_pool;
  p.knock = g_knock;

  auto kern = k_conv_pl<CLOG, J, GATE, SKIP, NST, NACC, NE, NI>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(conv_pl): %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  int grid = p.ntiles < 148 ? p.ntiles : 148;
  if (g_cap > 0 && grid > g_cap) grid = g_cap;
  launch_pdl(kern, dim3(grid), dim3(CF::WARPS * 32), (size_t)CF::SMEM, stream, p);
  count_launch();
  return check_launch("k_conv_pl");
}

}  // namespace pl

int* tc_status_word();   // conv_tc.cu: lazily allocated device status word

// Returns true if the planes kernel takes this launch (and runs it unless dry_run: rc = its return code).
bool conv_pl_try(const fn_conv_desc& d, cudaStream_t stream, bool dry_run, int& rc) {
  using namespace pl;
  rc = 0;
  if (g_enable < 0) {
    const char* e = getenv("FLOWNET_PL");
    g_enable = (e && e[0] == '0') ? 0 : 1;
  }
  if (!g_enable || !d.x_planes) return false;
  if (d.ksize != 3 || d.stride != 1 || d.prologue != FN_PRO_CONCAT_ELU || d.w_tc == nullptr) return false;
  if (d.keep_prob < 1.0f) return false;
  const bool gate = d.epilogue == FN_EPI_GATE_RES;
  if (!gate && d.epilogue != FN_EPI_BIAS) return false;
  if (d.Cout != (gate ? 2 * d.Cin : d.Cin)) return false;
  const bool skip = d.skip != nullptr;
  if (skip && (gate || d.Cskip != d.Cin || !d.skip_planes)) return false;
  if (d.y == nullptr && d.y_planes == nullptr) return false;
  if (get_encode() == nullptr) return false;
  const int clog = d.Cin == 8 ? 0 : d.Cin == 16 ? 1 : d.Cin == 32 ? 2 : -1;
  if (clog < 0) return false;
#define PL_CASE(CL, JJ, G, S, NS, NA, NEW, NIW, DNE, DNI)                                      \
  if (clog == CL && gate == G && skip == S && (g_ne ? g_ne : DNE) == NEW && (g_ni ? g_ni : DNI) == NIW && \
      (g_j == 0 || g_j == JJ)) {                                                                \
    if (!dry_run) {                                                                             \
      int* st = tc_status_word();                                                               \
      rc = !st ? (int)cudaErrorMemoryAllocation : launch<CL, JJ, G, S, NS, NA, NEW, NIW>(d, st, stream); \
    }                                                                                           \
    return true;                                                                                \
  }
  // every shape with 8 / 16 epilogue warps x 1 / 2 / 4 issuer warps (tools/ab_planes.py sweeps them); DNE / DNI = the
  // measured best at batch 64 (profiles/r2_ab_planes.log)
#define PL_SHAPE(CL, JJ, G, S, NS, NA, DNE, DNI)                                                   \
  PL_CASE(CL, JJ, G, S, NS, NA, 8, 1, DNE, DNI) PL_CASE(CL, JJ, G, S, NS, NA, 8, 2, DNE, DNI) PL_CASE(CL, JJ, G, S, NS, NA, 8, 4, DNE, DNI) \
  PL_CASE(CL, JJ, G, S, NS, NA, 16, 1, DNE, DNI) PL_CASE(CL, JJ, G, S, NS, NA, 16, 2, DNE, DNI) PL_CASE(CL, JJ, G, S, NS, NA, 16, 4, DNE, DNI)
  //        C  J  gate   skip   ring acc  warps issuers
  PL_SHAPE(0, 4, true,  false, 8, 8, 16, 2)
  PL_SHAPE(0, 4, false, false, 8, 8, 16, 4)
  PL_SHAPE(0, 4, false, true,  4, 8, 8, 2)
  PL_SHAPE(1, 2, true,  false, 8, 8, 8, 4)
  PL_SHAPE(1, 2, false, false, 8, 8, 8, 2)
  PL_SHAPE(1, 2, false, true,  4, 8, 8, 2)
  PL_SHAPE(1, 4, true,  false, 4, 4, 8, 1)
  PL_SHAPE(1, 4, false, false, 4, 8, 8, 2)
  PL_SHAPE(2, 1, true,  false, 4, 8, 8, 1)
  PL_SHAPE(2, 1, false, false, 8, 8, 8, 4)
  PL_SHAPE(2, 1, false, true,  4, 8, 8, 2)
  PL_SHAPE(2, 2, false, false, 4, 8, 8, 2)
#undef PL_SHAPE
#undef PL_CASE
  return false;
}

}  // namespace fn

extern "C" {
// test / timing hooks
void fn_debug_set_pl(int v) { fn::pl::g_enable = v ? 1 : 0; }
void fn_debug_set_pl_grid_cap(int v) { fn::pl::g_cap = v; }
void fn_debug_set_pl_ne(int v) { fn::pl::g_ne = v; }
void fn_debug_set_pl_ni(int v) { fn::pl::g_ni = v; }
void fn_debug_set_pl_j(int v) { fn::pl::g_j = v; }
void fn_debug_set_pl_knock(int v) { fn::pl::g_knock = v; }
}
