// Persistent, template-specialised tcgen05 kernel for the stride-1 3x3 convolutions of the
// high-resolution levels of flow_net (concat_elu prologue, C in {8,16,32}); these launches are
// HBM- and instruction-issue-bound, so everything that can be a compile-time constant is one.
//
// Differences from the generic k_conv_tc (conv_tc.cu), same math and same A/B layouts:
//   * persistent CTAs: grid = min(tiles, 148 * CTAs/SM); weights (<= 73 KB) are loaded ONCE per CTA by
//     the TMA engine and stay resident; TMEM is allocated once
//   * the raw halo tile, the skip tile and the shortcut tile arrive by TMA TENSOR loads
//     (cp.async.bulk.tensor.4d, out-of-image pixels zero-filled by the hardware = TF 'SAME' padding),
//     double-buffered so tile t+1 is in flight while tile t is transformed / multiplied / stored
//   * transform: shared -> registers -> shared, concat_elu with one ex2 per element, straight-line
//   * MMA issue: one elected lane per sub-tile, fully unrolled k-step program with immediate offsets
//   * epilogue: compile-time channel counts, shortcut read from the staged tile
// reference semantics: res_block conv_1 (+nin) and conv_2 (+gate, +shortcut), flow_architecture.py:129-165
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"

namespace fn {

constexpr int SP_THREADS = 160;   // warps 0-3: transform / issue / epilogue; warp 4: TMA producer + TMEM owner
constexpr int SP_HDR = 128 + 1024;

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
  static constexpr int PPAD = CH == 1 ? P : (P + ((CH >= 8 ? 1 : (CH == 4 ? 2 : 4)) - (P & 7) + 8) % 8);
  static constexpr int W_BYTES = KSTEPS * N * 32;
  static constexpr int PLANE_BYTES = 2 * CH * PPAD * 16;          // one operand, 2C channels
  static constexpr int STAGE_BYTES = ((P * C * 2) + 127) / 128 * 128;
  static constexpr int RES_BYTES = GATE ? 128 * J * F * 2 : 0;
  static constexpr int OFF_W = SP_HDR;
  static constexpr int OFF_PLANES = OFF_W + (W_BYTES + 127) / 128 * 128;
  static constexpr int OFF_PLANES_S = OFF_PLANES + (PLANE_BYTES + 127) / 128 * 128;
  static constexpr int OFF_STAGE = OFF_PLANES_S + (SKIP ? (PLANE_BYTES + 127) / 128 * 128 : 0);
  static constexpr int OFF_STAGE_S = OFF_STAGE + 2 * STAGE_BYTES;
  static constexpr int OFF_RES = OFF_STAGE_S + (SKIP ? 2 * STAGE_BYTES : 0);
  static constexpr int SMEM = OFF_RES + RES_BYTES + 128;           // +128: base alignment slack
  static constexpr int TMEM_COLS = (J * N <= 32) ? 32 : (J * N <= 64) ? 64 : (J * N <= 128) ? 128 : (J * N <= 256) ? 256 : 512;
};

// barriers (8 B each) inside the 128-byte header
enum { SB_W = 0, SB_X0 = 1, SB_X1 = 2, SB_S0 = 3, SB_S1 = 4, SB_R = 5, SB_DONE = 6, SB_COUNT = 7 };

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// stage [pp][C] raw bf16  ->  planes [2C/8][PPAD][8] = [elu(x) | elu(-x)], both bf16
template <typename Cfg>
__device__ __forceinline__ void transform_tile(const unsigned char* __restrict__ stage, unsigned char* __restrict__ planes,
                                               int tid) {
  constexpr int ITEMS = Cfg::P * Cfg::CH;
#pragma unroll 4
  for (int idx = tid; idx < ITEMS; idx += 128) {
    const int ch = idx & (Cfg::CH - 1);
    const int pp = idx >> (Cfg::CH == 1 ? 0 : (Cfg::CH == 2 ? 1 : (Cfg::CH == 4 ? 2 : 3)));
    bf16x8 in, o0, o1;
    in.v = *reinterpret_cast<const uint4*>(stage + (size_t)idx * 16);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v = __bfloat162float(in.h[j]);
      const float t = __expf(-fabsf(v)) - 1.f;
      o0.h[j] = __float2bfloat16(v > 0.f ? v : t);
      o1.h[j] = __float2bfloat16(v > 0.f ? t : -v);
    }
    unsigned char* d0 = planes + ((size_t)ch * Cfg::PPAD + pp) * 16;
    *reinterpret_cast<uint4*>(d0) = o0.v;
    *reinterpret_cast<uint4*>(d0 + (size_t)Cfg::CH * Cfg::PPAD * 16) = o1.v;
  }
}

// generic shortcut (avg-pool and/or front channel pad), 8 channels, global loads; reference :153-165
__device__ __forceinline__ void sp_shortcut_generic(const SpParams& p, int F, int b, int oy, int ox, int c0, float (&v)[8]) {
  const int shift = F - p.Cres;
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

template <int CLOG, int J, bool GATE, bool SKIP>
__global__ void __launch_bounds__(SP_THREADS) k_conv_s1p(const __grid_constant__ SpParams p) {
  using Cfg = SpCfg<CLOG, J, GATE, SKIP>;
  extern __shared__ unsigned char smem_raw[];
  // TMA destinations want 128-byte alignment
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  const uint32_t sbase = ptx::smem_u32(smem);
  auto bar = [&](int i) { return sbase + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + 64);
  float* sbias = reinterpret_cast<float*>(smem + 128);
  unsigned char* planes = smem + Cfg::OFF_PLANES;
  unsigned char* planes_s = smem + Cfg::OFF_PLANES_S;

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
  auto issue_tile_loads = [&](int t, int buf) {   // producer lane only
    int b, y0, x0;
    tile_coords(t, b, y0, x0);
    ptx::mbar_arrive_expect_tx(bar(SB_X0 + buf), (uint32_t)(Cfg::P * Cfg::C * 2));
    tma_load_4d(sbase + Cfg::OFF_STAGE + buf * Cfg::STAGE_BYTES, &p.tm_x, bar(SB_X0 + buf), 0, x0 - 1, y0 - 1, b);
    if (SKIP) {
      ptx::mbar_arrive_expect_tx(bar(SB_S0 + buf), (uint32_t)(Cfg::P * Cfg::C * 2));
      tma_load_4d(sbase + Cfg::OFF_STAGE_S + buf * Cfg::STAGE_BYTES, &p.tm_skip, bar(SB_S0 + buf), 0, x0 - 1, y0 - 1, b);
    }
  };

  // ---- one-time setup
  if (warp == 4) {
    ptx::tmem_alloc(sbase + 64, (uint32_t)Cfg::TMEM_COLS);
    ptx::tmem_relinquish();
    if (lane == 0) {
      for (int i = 0; i < SB_COUNT; ++i) ptx::mbar_init(bar(i), i == SB_DONE ? (uint32_t)J : 1u);
      ptx::fence_mbar_init();
      ptx::mbar_arrive_expect_tx(bar(SB_W), (uint32_t)Cfg::W_BYTES);
      ptx::bulk_g2s(sbase + Cfg::OFF_W, p.wpk, (uint32_t)Cfg::W_BYTES, bar(SB_W));
      if ((int)blockIdx.x < p.ntiles) issue_tile_loads(blockIdx.x, 0);
    }
    __syncwarp();
  }
  for (int i = tid; i < Cfg::N; i += SP_THREADS) sbias[i] = i < p.Cout ? __ldg(p.bias + i) : 0.f;
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t leader = ptx::elect_one();
  bool ok = true;

  int it = 0;
  for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++it) {
    const int buf = it & 1;
    const uint32_t par_buf = (uint32_t)((it >> 1) & 1), par_it = (uint32_t)(it & 1);
    int b, y0, x0;
    tile_coords(t, b, y0, x0);

    if (warp == 4) {
      // ---- producer: prefetch the next tile into the other buffer; fetch this tile's shortcut
      if (lane == 0) {
        const int tn = t + gridDim.x;
        if (tn < p.ntiles) issue_tile_loads(tn, buf ^ 1);
        if (GATE && p.res_mode == 1) {
          ptx::mbar_arrive_expect_tx(bar(SB_R), (uint32_t)Cfg::RES_BYTES);
          tma_load_4d(sbase + Cfg::OFF_RES, &p.tm_res, bar(SB_R), 0, x0, y0, b);
        }
      }
      __syncwarp();
    } else {
      // ---- transform: raw stage -> concat_elu planes
      ok = ok && ptx::mbar_wait(bar(SB_X0 + buf), par_buf);
      if (!ok) atomicCAS(p.status, 0, 11);
      transform_tile<Cfg>(smem + Cfg::OFF_STAGE + buf * Cfg::STAGE_BYTES, planes, tid);
      if (SKIP) {
        ok = ok && ptx::mbar_wait(bar(SB_S0 + buf), par_buf);
        if (!ok) atomicCAS(p.status, 0, 12);
        transform_tile<Cfg>(smem + Cfg::OFF_STAGE_S + buf * Cfg::STAGE_BYTES, planes_s, tid);
      }
      ptx::fence_proxy_async_smem();
    }
    ptx::tc_fence_before_sync();
    __syncthreads();                       // planes complete; stage[buf] consumed
    ptx::tc_fence_after_sync();

    if (warp < J) {
      // ---- MMA issue: sub-tile j = warp, fully unrolled program, immediate descriptor offsets
      if (it == 0) {
        ok = ok && ptx::mbar_wait(bar(SB_W), 0);
        if (!ok) atomicCAS(p.status, 0, 13);
        ptx::tc_fence_after_sync();
      }
      constexpr uint32_t idesc = ptx::idesc_bf16_f32(128, Cfg::N);
      const uint64_t a_tmpl = ptx::smem_desc(sbase + Cfg::OFF_PLANES, Cfg::PPAD * 16, Cfg::PU * 16) + (uint64_t)(8 * warp);
      const uint64_t b_tmpl = ptx::smem_desc(sbase + Cfg::OFF_W, Cfg::N * 16, 128);
      const uint32_t d_tmem = tmem_base + (uint32_t)(warp * Cfg::N);
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
        for (int kk = 0; kk < Cfg::KS; ++kk) {
          const uint32_t a_off = (uint32_t)(((tap / 3) * Cfg::PU + (tap % 3)) * 16 + kk * 2 * Cfg::PPAD * 16) >> 4;
          const uint32_t b_off = (uint32_t)((tap * Cfg::KS + kk) * Cfg::N * 32) >> 4;
          if (leader) ptx::mma_bf16_ss(d_tmem, a_tmpl + a_off, b_tmpl + b_off, idesc, (tap | kk) ? 1u : 0u);
        }
      }
      if (SKIP) {
        const uint64_t s_tmpl = ptx::smem_desc(sbase + Cfg::OFF_PLANES_S + (Cfg::PU + 1) * 16, Cfg::PPAD * 16, Cfg::PU * 16) +
                                (uint64_t)(8 * warp);
#pragma unroll
        for (int kk = 0; kk < Cfg::KS; ++kk) {
          const uint32_t a_off = (uint32_t)(kk * 2 * Cfg::PPAD * 16) >> 4;
          const uint32_t b_off = (uint32_t)((9 * Cfg::KS + kk) * Cfg::N * 32) >> 4;
          if (leader) ptx::mma_bf16_ss(d_tmem, s_tmpl + a_off, b_tmpl + b_off, idesc, 1u);
        }
      }
      if (leader) ptx::mma_commit(bar(SB_DONE));
      __syncwarp();
    }

    if (warp < 4) {
      // ---- epilogue: TMEM -> registers -> (+bias, gate, shortcut) -> bf16 NHWC
      ok = ok && ptx::mbar_wait(bar(SB_DONE), par_it);
      if (!ok) atomicCAS(p.status, 0, 14);
      ptx::tc_fence_after_sync();
      if (GATE && p.res_mode == 1) {
        ok = ok && ptx::mbar_wait(bar(SB_R), par_it);
        if (!ok) atomicCAS(p.status, 0, 15);
      }
      const int v = tid >> 3, u = tid & 7;
      const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
      const int oy = y0 + v;
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const int ox = x0 + 8 * j + u;
        const bool live = ok && oy < p.H && ox < p.W;
        const size_t pix = ((size_t)b * p.H + oy) * p.W + ox;
        const uint32_t col0 = lane_base + (uint32_t)(j * Cfg::N);
#pragma unroll
        for (int c0 = 0; c0 < Cfg::F; c0 += 8) {
          float o[8];
          if (GATE) {
            float a[8], g[8];
            ptx::tmem_ld8x2(col0 + c0, col0 + Cfg::F + c0, a, g);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = (a[i] + sbias[c0 + i]) * sigmoid_f(g[i] + sbias[Cfg::F + c0 + i]);
            if (p.res_mode == 1) {
              bf16x8 rr;
              rr.v = *reinterpret_cast<const uint4*>(smem + Cfg::OFF_RES + ((size_t)(v * 8 * J + 8 * j + u) * Cfg::F + c0) * 2);
#pragma unroll
              for (int i = 0; i < 8; ++i) o[i] += __bfloat162float(rr.h[i]);
            } else if (p.res_mode == 2 && live) {
              sp_shortcut_generic(p, Cfg::F, b, oy, ox, c0, o);
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
    ptx::tc_fence_before_sync();
    __syncthreads();                       // accumulators drained, planes and shortcut tile reusable
    ptx::tc_fence_after_sync();
  }

  if (warp == 4) ptx::tmem_dealloc(tmem_base, (uint32_t)Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

// bf16 NHWC tensor [B,H,W,C] with a box of (C, bw, bh, 1)
static bool make_map_nhwc(CUtensorMap* tm, const void* base, int B, int H, int W, int C, int bw, int bh) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  const cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)bw, (cuuint32_t)bh, 1u};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int* tc_status_word();   // conv_tc.cu: lazily allocated device status word

template <int CLOG, int J, bool GATE, bool SKIP>
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
      p.res_mode = 2;
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

  auto kern = k_conv_s1p<CLOG, J, GATE, SKIP>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(s1p): %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  int per_sm = (227 * 1024) / (Cfg::SMEM + 1024);
  const int tmem_lim = 512 / Cfg::TMEM_COLS;
  if (per_sm > tmem_lim) per_sm = tmem_lim;
  if (per_sm > 6) per_sm = 6;
  if (per_sm < 1) per_sm = 1;
  int grid = 148 * per_sm;
  if (grid > p.ntiles) grid = p.ntiles;
  kern<<<grid, SP_THREADS, Cfg::SMEM, stream>>>(p);
  count_launch();
  return check_launch("k_conv_s1p");
}

// Returns true if a specialised kernel exists for this launch (and runs it: rc = its return code).
bool conv_s1p_try(const fn_conv_desc& d, cudaStream_t stream, bool dry_run, int& rc) {
  rc = 0;
  if (d.ksize != 3 || d.stride != 1 || d.prologue != FN_PRO_CONCAT_ELU || d.keep_prob < 1.0f || d.w_tc == nullptr) return false;
  const bool gate = d.epilogue == FN_EPI_GATE_RES;
  if (!gate && d.epilogue != FN_EPI_BIAS) return false;
  if (d.Cout != (gate ? 2 * d.Cin : d.Cin)) return false;
  const bool skip = d.skip != nullptr;
  if (skip && (gate || d.Cskip != d.Cin)) return false;
  if (d.H < 16) return false;                       // short images use the transposed orientation of k_conv_tc
  if (get_encode() == nullptr) return false;
  const int clog = d.Cin == 8 ? 0 : d.Cin == 16 ? 1 : -1;
  if (clog < 0) return false;
  // widest tile that divides the row and still leaves room for >= 3 CTAs per SM (cases are ordered by J)
#define SP_CASE(CL, JJ, G, S)                                                                              \
  if (clog == CL && ((d.W % (8 * JJ)) == 0 || JJ == 1) && gate == G && skip == S && SpCfg<CL, JJ, G, S>::SMEM <= 75 * 1024) { \
    if (!dry_run) rc = launch_s1p<CL, JJ, G, S>(d, stream);                                                \
    return true;                                                                                           \
  }
  SP_CASE(0, 4, true, false) SP_CASE(0, 4, false, false) SP_CASE(0, 4, false, true)
  SP_CASE(0, 2, true, false) SP_CASE(0, 2, false, false) SP_CASE(0, 2, false, true)
  SP_CASE(0, 1, true, false) SP_CASE(0, 1, false, false) SP_CASE(0, 1, false, true)
  SP_CASE(1, 4, true, false) SP_CASE(1, 4, false, false) SP_CASE(1, 4, false, true)
  SP_CASE(1, 2, true, false) SP_CASE(1, 2, false, false) SP_CASE(1, 2, false, true)
  SP_CASE(1, 1, true, false) SP_CASE(1, 1, false, false) SP_CASE(1, 1, false, true)
#undef SP_CASE
  return false;
}

}  // namespace fn
