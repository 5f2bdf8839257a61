// tcgen05 kernel for the WIDEST stride-1 3x3 convolutions: C = 256 input channels (K = 512 per tap after concat_elu),
// Cout = 256 or 512 (gated) on 8-row images -- the 8x16 level of the filter_size = 16 variant (BASELINE configs[4],
// SURVEY F8), which exceeds both the shared memory (64 operand planes) and the N <= 256 limit of conv_s1d.cu.
//
// Same organisation as k_conv_s1d (warp 8 = producer: TMA tiles + weight ring; warps 0-7 = transform, issue, epilogue;
// transposed tile: one 8 x 16 image region = one M = 128 tile), with two changes:
//   * the contraction runs in KP = 2 PASSES of 128 raw channels: stage 128 channels -> 32 concat_elu planes -> the
//     k-steps of those planes for all 9 taps -> (MMAs done) -> next 128 channels into the same planes, all passes
//     accumulating into the same TMEM columns.  The packed weights are pass-major (the host packs the two K halves
//     separately and concatenates them).
//   * N = 512 is two MMAs of N = 256 per k-step (u columns, v columns), 512 TMEM columns.
// VAR_SKIP reuses the pass mechanism for conv_1 + nin skip at C = 128 (two 32-plane operands do not fit either): pass 0 =
// the 3x3 taps over x, pass 1 = the centre tap over the skip tensor `a` (x_1 += nin(concat_elu(a)), :136-142); the
// standard packed image [main k-steps | skip k-steps] is already pass-major.
// reference semantics: res_block conv_1 / conv_2 (+gate, +shortcut), flow_architecture.py:129-165.
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tma.cuh"

namespace fn {
using namespace tma;
namespace s1w {

constexpr int THREADS = 288;
constexpr int HDR = 256 + 2048;                    // mbarriers, TMEM slot, 512 bias floats
constexpr int CHUNK_BYTES = 32768;
constexpr int NST = 2;                             // weight ring stages
constexpr int KP = 2;                              // passes over the input channels
constexpr int CP = 128, CHP = CP / 8;              // raw channels / 8-channel chunks per pass
enum { VAR_WIDE = 0, VAR_SKIP = 1 };

struct Params {
  CUtensorMap tm_x, tm_res, tm_skip;
  CUtensorMap tm_w;                    // NSPLIT = 2: packed weights as 8-byte units (2 F, parts, 2 K halves, k-steps)
  const __nv_bfloat16* wpk;
  const float* bias;
  __nv_bfloat16* y;
  const __nv_bfloat16* res;
  int* status;
  int B, H, W;
  int tiles_u, tiles_v, ntiles;
  int Cout, res_mode, Cres, RH, RW, res_pool;
  int gate_bwd;
};

// NSPLIT = 2 (wide variant): two CTAs share a tile, each owns half of the output channels and of the weight stream
// (conv_s1d.cu has the same option; here a CTA's stream is 4.7 MB per tile without it)
template <bool GATE, int VAR, int NSPLIT = 1>
struct Cfg {
  static constexpr bool SKIPV = VAR == VAR_SKIP;
  static constexpr int C = SKIPV ? CP : KP * CP;                   // raw channels of x (the tensor-map extent)
  static constexpr int F = C;                                      // output channels
  static constexpr int N = GATE ? 2 * C : C;                       // 512 / 256 (wide), 128 (skip variant)
  static constexpr int FL = F / NSPLIT, NL = N / NSPLIT;           // output channels / accumulator columns of one CTA
  static constexpr int PARTS = GATE ? 2 : 1;                       // row blocks of F rows in the packed image (u, v)
  static constexpr int NMMA = NL <= 256 ? NL : 256;                // columns of one MMA
  static constexpr int NM = NL / NMMA;                             // MMAs per k-step
  static constexpr int KS = CHP;                                   // k-steps per tap and pass (2 * 128 / 16)
  static constexpr int KSTEPS_P0 = 9 * KS;                         // pass 0: all taps
  static constexpr int KSTEPS_P1 = SKIPV ? KS : 9 * KS;            // pass 1: the other channel half, or the skip's centre tap
  static constexpr int PU = 10, PV = 18, P = PU * PV;              // transposed tile: u = 8 rows (+halo) of y, v = 16 (+halo) of x
  static constexpr int PPAD = P + ((1 - (P & 7)) + 8) % 8;
  static constexpr int SPC = CHUNK_BYTES / (NL * 32);              // k-steps per weight chunk
  static constexpr int NCHUNKS_P0 = KSTEPS_P0 / SPC, NCHUNKS_P1 = KSTEPS_P1 / SPC;
  static constexpr int PLANE_BYTES = (2 * CHP * PPAD * 16 + 127) / 128 * 128;
  static constexpr int STAGE_BYTES = ((P * CP * 2) + 127) / 128 * 128;
  static constexpr int RES_BYTES = GATE ? 128 * F * 2 : 0;
  static constexpr int STAGE_REGION = STAGE_BYTES > RES_BYTES ? STAGE_BYTES : RES_BYTES;
  static constexpr int OFF_PLANES = HDR;
  static constexpr int OFF_STAGE = OFF_PLANES + PLANE_BYTES;
  static constexpr int OFF_RING = OFF_STAGE + STAGE_REGION;
  static constexpr int SMEM = OFF_RING + NST * CHUNK_BYTES + 128;
  static constexpr int TMEM_COLS = NL;                             // 128, 256 or 512 (powers of two)
  static_assert(KSTEPS_P0 % SPC == 0 && KSTEPS_P1 % SPC == 0, "a weight chunk must not straddle two passes");
  static_assert(!(SKIPV && GATE), "the skip operand belongs to conv_1 (no gate)");
  static_assert(NSPLIT == 1 || (!SKIPV && 2 * FL <= 256), "channel split: wide variant, box of at most 256 8-byte units");
  static_assert(SMEM <= 227 * 1024, "shared memory");
};

enum { SB_X = 0, SB_R = 1, SB_TF = 2, SB_TE = 3, SB_ST = 4, SB_PD = 5, SB_FULL = 6, SB_EMPTY = 8 };

__device__ __forceinline__ void worker_barrier() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// stage [pu][pv][CP] raw bf16 (TMA box order, transposed tile) -> planes [2 CHP][PPAD][8] = [elu | elu(-)], index pv*PU+pu
template <typename CF>
__device__ __forceinline__ void transform_tile(uint32_t stage, uint32_t planes, int tid) {
  constexpr int ITEMS = CF::P * CHP;
#pragma unroll 2
  for (int idx = tid; idx < ITEMS; idx += 256) {
    const int ch = idx & (CHP - 1);
    const int sp = idx >> 4;
    const int pu = sp / CF::PV, pv = sp - pu * CF::PV;
    const int pp = pv * CF::PU + pu;
    bf16x8 in, o0, o1;
    in.v = lds128(stage + (uint32_t)idx * 16u);
    concat_elu8(in.v, o0.v, o1.v);
    const uint32_t d0 = planes + (uint32_t)(ch * CF::PPAD + pp) * 16u;
    sts128(d0, o0.v);
    sts128(d0 + (uint32_t)(CHP * CF::PPAD * 16), o1.v);
  }
}

template <bool GATE, int VAR, int NSPLIT = 1>
__global__ void __launch_bounds__(THREADS) k_conv_s1w(const __grid_constant__ Params p) {
  using CF = Cfg<GATE, VAR, NSPLIT>;
  constexpr int F = CF::F;
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

  // tile -> (image, u0 along y, v0 along x)
  auto tile_coords = [&](int t, int& b, int& u0, int& v0) {
    const int tu = t % p.tiles_u;
    const int r = t / p.tiles_u;
    const int tv = r % p.tiles_v;
    b = r / p.tiles_v;
    u0 = tu * 8;
    v0 = tv * 16;
  };
  const int nh = (int)blockIdx.x % NSPLIT;                  // channel half of this CTA
  const int cta = (int)blockIdx.x / NSPLIT, ncta = (int)gridDim.x / NSPLIT;
  const int n_local = (cta < p.ntiles) ? (p.ntiles - 1 - cta) / ncta + 1 : 0;

  if (warp == 8) {
    ptx::tmem_alloc(sbase + 128, (uint32_t)CF::TMEM_COLS);
    ptx::tmem_relinquish();
    if (lane == 0) {
      ptx::mbar_init(bar(SB_X), 1);
      ptx::mbar_init(bar(SB_R), 1);
      ptx::mbar_init(bar(SB_TF), 1);
      ptx::mbar_init(bar(SB_TE), 256);
      ptx::mbar_init(bar(SB_ST), 256);
      ptx::mbar_init(bar(SB_PD), 1);
      for (int s = 0; s < NST; ++s) {
        ptx::mbar_init(bar(SB_FULL + s), 1);
        ptx::mbar_init(bar(SB_EMPTY + s), 1);
      }
      ptx::fence_mbar_init();
    }
    __syncwarp();
  }
  for (int i = tid; i < CF::N; i += THREADS) {
    float bv = i < p.Cout ? __ldg(p.bias + i) : 0.f;
    if (GATE && i >= F) bv *= 0.5f;
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
      int gchunk = 0;
      for (int it = 0; it < n_local && ok; ++it) {
        int b, u0, v0;
        tile_coords(cta + it * ncta, b, u0, v0);
        for (int ps = 0; ps < KP && ok; ++ps) {
          const int q = it * KP + ps;                   // running (tile, pass) index: parities of X and ST
          if (ps == 0 && it >= 1) {                     // stage region (aliased by the shortcut tile) free: epilogue done
            ok = ptx::mbar_wait(bar(SB_TE), (uint32_t)((it - 1) & 1));
            if (!ok) { atomicCAS(p.status, 0, 71); break; }
          } else if (ps >= 1) {                         // stage consumed by the previous pass's transform
            ok = ptx::mbar_wait(bar(SB_ST), (uint32_t)((q - 1) & 1));
            if (!ok) { atomicCAS(p.status, 0, 72); break; }
          }
          ptx::mbar_arrive_expect_tx(bar(SB_X), (uint32_t)(CF::P * CP * 2));
          if (CF::SKIPV && ps == 1) tma_load_4d(sbase + CF::OFF_STAGE, &p.tm_skip, bar(SB_X), 0, v0 - 1, u0 - 1, b);
          else tma_load_4d(sbase + CF::OFF_STAGE, &p.tm_x, bar(SB_X), ps * CP, v0 - 1, u0 - 1, b);
          const int nchunks = ps == 0 ? CF::NCHUNKS_P0 : CF::NCHUNKS_P1;
          for (int c = 0; c < nchunks && ok; ++c, ++gchunk) {
            const int s = gchunk % NST;
            if (gchunk >= NST) {
              ok = ptx::mbar_wait(bar(SB_EMPTY + s), (uint32_t)(((gchunk / NST) - 1) & 1));
              if (!ok) { atomicCAS(p.status, 0, 73); break; }
            }
            const uint32_t bytes = (uint32_t)(CF::SPC * CF::NL * 32);
            ptx::mbar_arrive_expect_tx(bar(SB_FULL + s), bytes);
            if (NSPLIT == 1)
              ptx::bulk_g2s(sbase + CF::OFF_RING + s * CHUNK_BYTES,
                            reinterpret_cast<const unsigned char*>(p.wpk) + (size_t)(ps * CF::NCHUNKS_P0 + c) * CF::SPC * CF::N * 32, bytes,
                            bar(SB_FULL + s));
            else       // this CTA's rows of the u (and v) block, both K halves, SPC k-steps: [kstep][khalf][u rows | v rows]
              tma_load_4d(sbase + CF::OFF_RING + s * CHUNK_BYTES, &p.tm_w, bar(SB_FULL + s), 2 * nh * CF::FL, 0, 0,
                          (ps * CF::NCHUNKS_P0 + c) * CF::SPC);
          }
        }
        if (ok && GATE && p.res_mode == 1) {            // last pass's transform has consumed the stage: fetch the shortcut tile
          ok = ptx::mbar_wait(bar(SB_ST), (uint32_t)((it * KP + KP - 1) & 1));
          if (!ok) { atomicCAS(p.status, 0, 74); break; }
          ptx::mbar_arrive_expect_tx(bar(SB_R), (uint32_t)CF::RES_BYTES);
          tma_load_4d(sbase + CF::OFF_STAGE, &p.tm_res, bar(SB_R), 0, v0, u0, b);
        }
      }
    }
    __syncwarp();
  } else {
    // =========================== workers ===========================
    const uint32_t leader = ptx::elect_one();
    constexpr uint32_t idesc = ptx::idesc_bf16_f32(128, CF::NMMA);
    const int q4 = warp & 3, half = warp >> 2;
    const int r = q4 * 32 + lane;                     // accumulator row == TMEM lane
    const int v = r >> 3, u = r & 7;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q4 * 32) << 16);
    int gchunk = 0;
    for (int it = 0; it < n_local; ++it) {
      const uint32_t par_it = (uint32_t)(it & 1);
      int b, u0, v0;
      tile_coords(cta + it * ncta, b, u0, v0);
      for (int ps = 0; ps < KP; ++ps) {
        const int q = it * KP + ps;
        ok = ok && ptx::mbar_wait(bar(SB_X), (uint32_t)(q & 1));
        if (ps >= 1) {                                // MMAs of the previous pass no longer read the planes
          const int d = it * (KP - 1) + ps - 1;
          ok = ok && ptx::mbar_wait(bar(SB_PD), (uint32_t)(d & 1));
        }
        if (!ok) atomicCAS(p.status, 0, 75);
        transform_tile<CF>(sbase + CF::OFF_STAGE, sbase + CF::OFF_PLANES, tid);
        ptx::mbar_arrive(bar(SB_ST));
        ptx::fence_proxy_async_smem();
        worker_barrier();
        if (warp == 0) {
          ptx::tc_fence_after_sync();
          const uint64_t a_tmpl = ptx::smem_desc(sbase + CF::OFF_PLANES, CF::PPAD * 16, CF::PU * 16);
          const uint64_t b_tmpl = ptx::smem_desc(sbase + CF::OFF_RING, CF::NL * 16, 128);
          const int nchunks = ps == 0 ? CF::NCHUNKS_P0 : CF::NCHUNKS_P1;
          const bool centre_only = CF::SKIPV && ps == 1;
#pragma unroll 1
          for (int c = 0; c < nchunks; ++c) {
            const int gc = gchunk + c;
            const int s = gc % NST;
            ok = ok && ptx::mbar_wait(bar(SB_FULL + s), (uint32_t)((gc / NST) & 1));
            if (!ok) atomicCAS(p.status, 0, 76);
            ptx::tc_fence_after_sync();
            const uint64_t b_chunk = b_tmpl + (uint64_t)((uint32_t)(s * CHUNK_BYTES) >> 4);
#pragma unroll
            for (int i = 0; i < CF::SPC; ++i) {
              const int ks = c * CF::SPC + i;                       // k-step inside this pass
              const int tap = centre_only ? 4 : ks / CF::KS, kk = ks % CF::KS;
              const int di = tap / 3, dj = tap % 3;
              const int dv = dj, du = di;                           // transposed tile: u along y
              const uint32_t a_off = (uint32_t)((dv * CF::PU + du) * 16 + kk * 2 * CF::PPAD * 16) >> 4;
              const uint32_t b_off = (uint32_t)(i * CF::NL * 32) >> 4;
              const uint32_t acc = (ps == 0 && ks == 0) ? 0u : 1u;
#pragma unroll
              for (int m = 0; m < CF::NM; ++m)
                if (leader)
                  ptx::mma_bf16_ss(tmem_base + (uint32_t)(m * CF::NMMA), a_tmpl + a_off, b_chunk + b_off + (uint32_t)((m * CF::NMMA * 16) >> 4), idesc,
                                   acc);
            }
            if (leader) ptx::mma_commit(bar(SB_EMPTY + s));
          }
          if (leader) ptx::mma_commit(bar(ps == KP - 1 ? SB_TF : SB_PD));
          __syncwarp();
        }
        gchunk += ps == 0 ? CF::NCHUNKS_P0 : CF::NCHUNKS_P1;
      }
      // ---- epilogue: quadrant q4, channel half `half`
      ok = ok && ptx::mbar_wait(bar(SB_TF), par_it);
      if (!ok) atomicCAS(p.status, 0, 77);
      ptx::tc_fence_after_sync();
      if (GATE && p.res_mode == 1) {
        ok = ok && ptx::mbar_wait(bar(SB_R), par_it);
        if (!ok) atomicCAS(p.status, 0, 78);
      }
      {
        const int oy = u0 + u, ox = v0 + v;
        const bool live = ok && oy < p.H && ox < p.W;
        const size_t pix = ((size_t)b * p.H + oy) * p.W + ox;
        const uint32_t res_pix = (uint32_t)(u * 16 + v);          // shortcut tile = TMA box order [8 of u][16 of v]
#pragma unroll 2
        for (int g8 = 0; g8 < CF::FL / 16; ++g8) {
          const int cl = half * (CF::FL / 2) + g8 * 8;      // column inside this CTA's accumulator
          const int c0 = nh * CF::FL + cl;                  // output channel
          float o[8];
          if (GATE) {
            float a[8], g[8];
            ptx::tmem_ld8x2(lane_base + cl, lane_base + CF::FL + cl, a, g);
            if (p.gate_bwd) {
              bf16x8 gy, du, dv;
              gy.v = lds128(sbase + CF::OFF_STAGE + (res_pix * F + (uint32_t)c0) * 2u);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float sg = fmaf(tanh_approx(fmaf(g[i], 0.5f, sbias[F + c0 + i])), 0.5f, 0.5f);
                const float gs = __bfloat162float(gy.h[i]) * sg;
                du.h[i] = __float2bfloat16(gs);
                dv.h[i] = __float2bfloat16(gs * (a[i] + sbias[c0 + i]) * (1.f - sg));
              }
              if (live) {
                *reinterpret_cast<uint4*>(p.y + pix * (2 * F) + c0) = du.v;
                *reinterpret_cast<uint4*>(p.y + pix * (2 * F) + F + c0) = dv.v;
              }
              continue;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float th = tanh_approx(fmaf(g[i], 0.5f, sbias[F + c0 + i]));
              o[i] = (a[i] + sbias[c0 + i]) * fmaf(th, 0.5f, 0.5f);
            }
            if (p.res_mode == 1) {
              bf16x8 rr;
              rr.v = lds128(sbase + CF::OFF_STAGE + (res_pix * F + (uint32_t)c0) * 2u);
#pragma unroll
              for (int i = 0; i < 8; ++i) o[i] += __bfloat162float(rr.h[i]);
            } else if (p.res_mode == 2 && live) {
              // generic shortcut (pooled / channel-padded): plain global loads
              const int shift = F - p.Cres;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int cr = c0 + i - shift;
                if (cr < 0) continue;
                if (p.res_pool) {
                  float s4 = 0.f;
                  int cnt = 0;
                  for (int dy = 0; dy < 2; ++dy)
                    for (int dx = 0; dx < 2; ++dx) {
                      const int ry = 2 * oy + dy, rx = 2 * ox + dx;
                      if (ry < p.RH && rx < p.RW) {
                        s4 += __bfloat162float(p.res[(((size_t)b * p.RH + ry) * p.RW + rx) * p.Cres + cr]);
                        ++cnt;
                      }
                    }
                  o[i] += s4 / (float)cnt;
                } else {
                  o[i] += __bfloat162float(p.res[(((size_t)b * p.RH + oy) * p.RW + ox) * p.Cres + cr]);
                }
              }
            }
          } else {
            ptx::tmem_ld8(lane_base + cl, o);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] += sbias[c0 + i];
          }
          if (live) {
            bf16x8 st;
#pragma unroll
            for (int i = 0; i < 8; ++i) st.h[i] = __float2bfloat16(o[i]);
            *reinterpret_cast<uint4*>(p.y + pix * F + c0) = st.v;
          }
        }
      }
      ptx::tc_fence_before_sync();
      ptx::mbar_arrive(bar(SB_TE));
      if (it + 1 < n_local) {
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

template <bool GATE, int VAR, int NSPLIT = 1>
static int launch(const fn_conv_desc& d, int* status, cudaStream_t stream) {
  using CF = Cfg<GATE, VAR, NSPLIT>;
  constexpr int C = CF::C, F = CF::F;
  Params p;
  memset(&p, 0, sizeof(p));
  // transposed tile: the box is (channels, x-extent = PV, y-extent = PU)
  PFN_encodeTiled enc = get_encode();
  auto map_cp = [&](CUtensorMap* tm, const void* base, int c_total, int H, int W) {
    const cuuint64_t dims[4] = {(cuuint64_t)c_total, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)d.B};
    const cuuint64_t strides[3] = {(cuuint64_t)c_total * 2, (cuuint64_t)W * c_total * 2, (cuuint64_t)H * W * c_total * 2};
    const cuuint32_t box[4] = {(cuuint32_t)CP, (cuuint32_t)CF::PV, (cuuint32_t)CF::PU, 1u};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc && enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  if (!map_cp(&p.tm_x, d.x, C, d.H, d.W)) { set_error("conv_s1w: tensor map (x) failed"); return FN_E_UNSUPPORTED; }
  if (CF::SKIPV && !map_cp(&p.tm_skip, d.skip, CP, d.SH, d.SW)) { set_error("conv_s1w: tensor map (skip) failed"); return FN_E_UNSUPPORTED; }
  p.res_mode = 0;
  if (GATE && d.res != nullptr) {
    if (!d.res_pool && d.Cres == F) {
      p.res_mode = 1;
      if (!make_map_nhwc(&p.tm_res, d.res, d.B, d.RH, d.RW, F, 16, 8)) { set_error("conv_s1w: tensor map (res) failed"); return FN_E_UNSUPPORTED; }
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
  p.tiles_u = (d.H + 7) / 8;
  p.tiles_v = (d.W + 15) / 16;
  p.ntiles = d.B * p.tiles_u * p.tiles_v;
  p.Cout = d.Cout;
  p.Cres = d.Cres; p.RH = d.RH; p.RW = d.RW; p.res_pool = d.res_pool;
  p.gate_bwd = d.epilogue == FN_EPI_GATE_BWD ? 1 : 0;
  if (NSPLIT > 1) {
    const cuuint64_t dims[4] = {(cuuint64_t)(2 * F), (cuuint64_t)CF::PARTS, 2, (cuuint64_t)(CF::KSTEPS_P0 + CF::KSTEPS_P1)};
    const cuuint64_t strides[3] = {(cuuint64_t)F * 16, (cuuint64_t)CF::N * 16, (cuuint64_t)CF::N * 32};
    const cuuint32_t box[4] = {(cuuint32_t)(2 * CF::FL), (cuuint32_t)CF::PARTS, 2, (cuuint32_t)CF::SPC};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    if (!enc || enc(&p.tm_w, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void*>(d.w_tc), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      set_error("conv_s1w: tensor map (weights) failed");
      return FN_E_UNSUPPORTED;
    }
  }
  auto kern = k_conv_s1w<GATE, VAR, NSPLIT>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(s1w): %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  const int grid = (p.ntiles < 148 / NSPLIT ? p.ntiles : 148 / NSPLIT) * NSPLIT;
  kern<<<grid, THREADS, CF::SMEM, stream>>>(p);
  count_launch();
  return check_launch("k_conv_s1w");
}

}  // namespace s1w

int* tc_status_word();

// C = 256 stride-1 3x3 concat_elu convs on images with H % 8 == 0 rows (the 8x16 level of the wide variant).
// d.w_tc must be the PASS-MAJOR image: fn_pack_conv_weights_tc of the K rows {0..127, 256..383} followed by that of the
// rows {128..255, 384..511} (flow_architecture._conv_weights does this for 512-row weights).
bool conv_s1w_try(const fn_conv_desc& d, cudaStream_t stream, bool dry_run, int& rc) {
  using namespace s1w;
  rc = 0;
  if (d.ksize != 3 || d.stride != 1 || d.prologue != FN_PRO_CONCAT_ELU || d.w_tc == nullptr || d.keep_prob < 1.0f) return false;
  if ((d.H % 8) != 0 || d.W < 16 || get_encode() == nullptr) return false;
  if (d.skip != nullptr) {
    // conv_1 + nin skip at C = 128: x pass + skip pass
    if (d.Cin != CP || d.Cskip != CP || d.Cout != CP || d.epilogue != FN_EPI_BIAS) return false;
    if (!dry_run) {
      int* st = tc_status_word();
      rc = !st ? (int)cudaErrorMemoryAllocation : launch<false, VAR_SKIP>(d, st, stream);
    }
    return true;
  }
  constexpr int C = KP * CP;
  if (d.Cin != C) return false;
  const bool gate = d.epilogue == FN_EPI_GATE_RES || d.epilogue == FN_EPI_GATE_BWD;
  if (!gate && d.epilogue != FN_EPI_BIAS) return false;
  if (d.Cout != (gate ? 2 * C : C)) return false;
  if (d.epilogue == FN_EPI_GATE_BWD && (d.res == nullptr || d.res_pool || d.Cres != C)) return false;
  if (!dry_run) {
    int* st = tc_status_word();
    const bool split = d.B * ((d.H + 7) / 8) * ((d.W + 15) / 16) <= 74;       // under half a wave of tiles: two CTAs per tile
    rc = !st ? (int)cudaErrorMemoryAllocation
             : split ? (gate ? launch<true, VAR_WIDE, 2>(d, st, stream) : launch<false, VAR_WIDE, 2>(d, st, stream))
                     : (gate ? launch<true, VAR_WIDE>(d, st, stream) : launch<false, VAR_WIDE>(d, st, stream));
  }
  return true;
}

}  // namespace fn
