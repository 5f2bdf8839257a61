// The tail of flow_net on the tensor cores: last_conv (3x3, 8 -> 2 channels, no prologue) + tanh, fp32 NHWC out
// (reference flow_architecture.py:242-243).
//
// With K = 8 per tap and N = 2 the layer looks nothing like a GEMM, and a first tensor-core version (one MMA per tap on
// the stride-1 kernel's skeleton, K zero-padded to 16) only matched the CUDA-core kernel: a tcgen05.mma occupies the MMA
// unit ~39 clocks whatever its N and K, nine of them per 128 pixels are 19.8 us at batch 64.  Two
// observations make it worthwhile:
//   * the input needs no prologue, so the raw NHWC halo tile a tensor copy delivers ([row][col][8 channels], 16 bytes per
//     pixel) IS the UMMA A operand in its no-swizzle K-major form -- no CUDA-core pass over the input at all;
//   * the two K halves of one MMA may sit anywhere in shared memory (leading-dimension byte offset of the matrix
//     descriptor), so ONE instruction takes TWO taps: K half 0 = the 8 channels under tap t, K half 1 = the same plane
//     shifted to tap t + 1 (LBO = the distance between the two taps).  Nine taps are five MMAs (the tenth half meets
//     zero weights): 11 us of MMA-unit time at batch 64 instead of 19.8.
// What is left for the CUDA cores is the epilogue: two accumulator columns per pixel, bias, tanh, one float2 store --
// against 144 FMAs + 18 shared loads per pixel in k_tail8_v2 (21.1 us back to back, 68 % of the issue slots); measured 16.9 us,
// of which 11 us are the MMA unit's.
//
// Organisation as conv_wg.cu: a CTA is NWG autonomous warpgroups; each owns two tile buffers (the tensor copy of tile
// i + 1 lands while tile i is multiplied and written out) and J accumulators of 16 columns.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"
#include "tma.cuh"

namespace fn {
using namespace tma;
namespace tail {

constexpr int HDR = 1024;          // [0,256) mbarriers, [256] TMEM base, [512, 520) bias
constexpr int KSTEPS = 5, N = 16;
constexpr int W_BYTES = KSTEPS * N * 32;

struct Params {
  CUtensorMap tm_x;
  const __nv_bfloat16* w;       // [9][8][2] (tap, channel, output)
  const float* bias;            // [2]
  float* y;                     // [B,H,W,2]
  int* status;
  int B, H, W;
  int tiles_u, tiles_v, ntiles;
};

template <int J, int NWG>
struct Cfg {
  static constexpr int PU = 8 * J + 2, PV = 18;
  static constexpr int TILE_BYTES = (PU * PV * 16 + 127) / 128 * 128;
  static constexpr int OFF_W = HDR;
  static constexpr int OFF_T = OFF_W + (W_BYTES + 127) / 128 * 128;
  static constexpr int SMEM = OFF_T + NWG * 2 * TILE_BYTES;
  static constexpr int ACC_COLS = J * N;
  static constexpr int ALL = NWG * ACC_COLS;
  static constexpr int TMEM_COLS = ALL <= 32 ? 32 : ALL <= 64 ? 64 : ALL <= 128 ? 128 : ALL <= 256 ? 256 : 512;
  static_assert(ALL <= 512 && J <= 4, "TMEM / one issuing warp per sub-tile");
  static_assert((1 + NWG * (2 + J)) * 8 <= 256, "mbarrier area");
};

__device__ __forceinline__ void wg_barrier(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, float& a, float& b) {
  uint32_t r0, r1;
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r0), "=r"(r1)
      : "r"(taddr));
  a = __uint_as_float(r0);
  b = __uint_as_float(r1);
}

template <int J, int NWG, int MINB>
__global__ void __launch_bounds__(NWG * 128, MINB) k_tail_tc(const __grid_constant__ Params p) {
  using CF = Cfg<J, NWG>;
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = ptx::smem_u32(smem);
  auto bar = [&](int i) { return sbase + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + 256);
  float* sbias = reinterpret_cast<float*>(smem + 512);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int g = warp >> 2, q = warp & 3;

  if (threadIdx.x == 64) prefetch_tensormap(&p.tm_x);
  if (warp == 0) {
    ptx::tmem_alloc(sbase + 256, (uint32_t)CF::TMEM_COLS);
    ptx::tmem_relinquish();
  } else if (warp == 1 && lane == 0) {
    for (int i = 0; i < NWG * (2 + J); ++i) ptx::mbar_init(bar(i), 1);
    ptx::fence_mbar_init();
  }
  // B operand, built in place from the [tap][channel][output] weights: k-step s, K half h = tap 2s + h, rows n < 2 live
  //   [s][h][n][8 channels]  (LBO = N * 16 between the halves, SBO = 128 between groups of eight rows)
  {
    __nv_bfloat16* wb = reinterpret_cast<__nv_bfloat16*>(smem + CF::OFF_W);
    for (int i = threadIdx.x; i < KSTEPS * 2 * N * 8; i += NWG * 128) {
      const int c = i & 7, n = (i >> 3) % N, sh = i / (8 * N);      // sh = 2 s + h = tap
      wb[i] = (n < 2 && sh < 9) ? p.w[(sh * 8 + c) * 2 + n] : __float2bfloat16(0.f);
    }
    if (threadIdx.x < 2) sbias[threadIdx.x] = p.bias[threadIdx.x];
  }
  ptx::fence_proxy_async_smem();        // the weights are read by the tensor core
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  // ---- this warpgroup's resources
  const uint32_t tile0 = sbase + CF::OFF_T + (uint32_t)(g * 2 * CF::TILE_BYTES);
  const uint32_t bar_in = bar(g * (2 + J)), bar_mma = bar_in + 16u;
  const uint32_t acc = tmem_base + (uint32_t)(g * CF::ACC_COLS);
  const int stride = (int)gridDim.x * NWG;
  int t = (int)blockIdx.x * NWG + g;
  auto coords = [&](int tt, int& b, int& y0, int& x0) {
    const int tu = tt % p.tiles_u, rest = tt / p.tiles_u;
    x0 = tu * 8 * J;
    y0 = (rest % p.tiles_v) * 16;
    b = rest / p.tiles_v;
  };
  auto issue_load = [&](int tt, int buf) {
    int b, y0, x0;
    coords(tt, b, y0, x0);
    ptx::mbar_arrive_expect_tx(bar_in + 8u * (uint32_t)buf, (uint32_t)(CF::PU * CF::PV * 16));
    tma_load_4d(tile0 + (uint32_t)(buf * CF::TILE_BYTES), &p.tm_x, bar_in + 8u * (uint32_t)buf, 0, x0 - 1, y0 - 1, b);
  };
  const uint32_t leader = (q < J) ? ptx::elect_one() : 0u;
  if (q == 0 && leader && t < p.ntiles) issue_load(t, 0);

  const int r = q * 32 + lane;             // accumulator row == TMEM lane
  const int v = r >> 3, u = r & 7;
  const uint32_t lane_acc = acc + ((uint32_t)(q * 32) << 16);
  constexpr uint32_t idesc = ptx::idesc_bf16_f32(128, N);
  const float b0 = sbias[0], b1 = sbias[1];
  bool ok = true;

  for (int i = 0; t < p.ntiles; t += stride, ++i) {
    const int buf = i & 1;
    int b, y0, x0;
    coords(t, b, y0, x0);
    ok = ok && ptx::mbar_wait(bar_in + 8u * (uint32_t)buf, (uint32_t)((i >> 1) & 1));
    if (!ok) atomicCAS(p.status, 0, 61);
    ptx::tc_fence_before_sync();            // this thread's TMEM reads of the previous tile are done
    wg_barrier(g);
    if (q < J) {
      ptx::tc_fence_after_sync();
      // the other buffer was read by the previous tile's MMAs, all complete (the epilogue waited for them)
      if (q == 0 && leader && t + stride < p.ntiles) issue_load(t + stride, buf ^ 1);
      __syncwarp();
      const uint32_t a0 = tile0 + (uint32_t)(buf * CF::TILE_BYTES) + (uint32_t)(q * 8 * 16);
      const uint64_t b_tmpl = ptx::smem_desc(sbase + CF::OFF_W, N * 16, 128);
      const uint32_t d_tmem = acc + (uint32_t)(q * N);
      if (leader) {
#pragma unroll
        for (int s = 0; s < KSTEPS; ++s) {
          // taps 2s and 2s + 1 as the two K halves of one instruction: LBO = the byte distance between their views
          const int t0 = 2 * s, t1 = 2 * s + 1 < 9 ? 2 * s + 1 : 2 * s;
          const uint32_t off0 = (uint32_t)(((t0 / 3) * CF::PU + (t0 % 3)) * 16), off1 = (uint32_t)(((t1 / 3) * CF::PU + (t1 % 3)) * 16);
          const uint32_t lbo = t1 > t0 ? off1 - off0 : 0u;      // the unpaired last tap: both halves read the same (finite) data, the second meets zero weights
          ptx::mma_bf16_ss(d_tmem, ptx::smem_desc(a0 + off0, lbo, CF::PU * 16), b_tmpl + (uint32_t)((s * N * 32) >> 4), idesc, s ? 1u : 0u);
        }
        ptx::mma_commit(bar_mma + 8u * (uint32_t)q);
      }
      __syncwarp();
    }
    // ---- epilogue: two accumulator columns per pixel -> bias, tanh -> float2
    const int oy = y0 + v;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int ox = x0 + 8 * j + u;
      ok = ok && ptx::mbar_wait(bar_mma + 8u * (uint32_t)j, (uint32_t)(i & 1));
      if (!ok) atomicCAS(p.status, 0, 62);
      ptx::tc_fence_after_sync();
      float a, c;
      tmem_ld2(lane_acc + (uint32_t)(j * N), a, c);
      if (ok && oy < p.H && ox < p.W)
        *reinterpret_cast<float2*>(p.y + (((size_t)b * p.H + oy) * p.W + ox) * 2) = make_float2(tanhf(a + b0), tanhf(c + b1));
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, (uint32_t)CF::TMEM_COLS);
  }
}

static int g_enable = 1;

template <int J, int NWG, int MINB>
static int launch(const fn_conv_desc& d, cudaStream_t stream, int* status) {
  using CF = Cfg<J, NWG>;
  Params p;
  if (!make_map_nhwc(&p.tm_x, d.x, d.B, d.H, d.W, 8, CF::PU, CF::PV)) { set_error("conv_tail: tensor map failed"); return FN_E_UNSUPPORTED; }
  p.w = (const __nv_bfloat16*)d.w;
  p.bias = d.bias;
  p.y = (float*)d.y;
  p.status = status;
  p.B = d.B; p.H = d.H; p.W = d.W;
  p.tiles_u = (d.W + 8 * J - 1) / (8 * J);
  p.tiles_v = (d.H + 15) / 16;
  p.ntiles = d.B * p.tiles_u * p.tiles_v;
  static bool configured = false;
  static int sms = 0;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_tail_tc<J, NWG, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM);
    int dev = 0;
    if (e == cudaSuccess) e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) { set_error("conv_tail: %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  const int want = (p.ntiles + NWG - 1) / NWG;
  const int grid = want < sms * MINB ? want : sms * MINB;
  k_tail_tc<J, NWG, MINB><<<grid, NWG * 128, CF::SMEM, stream>>>(p);
  count_launch();
  note_tcgen05(true);
  return check_launch("k_tail_tc");
}

}  // namespace tail

int* tc_status_word();   // conv_tc.cu

// last_conv + tanh at 8 input channels on the tensor cores; false: not this kernel's shape (the CUDA-core tail runs)
bool conv_tail_supported(const fn_conv_desc& d) {
  // measured at 128x256 (tools/ab_headtail.py): 21.1 -> 16.9 us at batch 64 (MMA-unit floor 11 us), 4.6 -> 5.9 us at batch 8 --
  // the persistent kernel needs a few tiles per warpgroup to pay for its set-up; FN_IMPL_TCGEN05 forces it
  const long long tiles = (long long)d.B * ((d.W + 31) / 32) * ((d.H + 15) / 16);
  if (d.impl != FN_IMPL_TCGEN05 && tiles < 3000) return false;
  return tail::g_enable && d.Cin == 8 && d.Cout == 2 && d.ksize == 3 && d.stride == 1 && d.prologue == FN_PRO_NONE && d.epilogue == FN_EPI_TANH &&
         !d.skip && !d.res && d.keep_prob >= 1.0f && !d.x_planes && d.y != nullptr && d.W >= 34 && d.H >= 18 && tma::get_encode() != nullptr;
}
int conv_tail_fwd(const fn_conv_desc& d, cudaStream_t stream) {
  int* status = tc_status_word();
  if (status == nullptr) return (int)cudaErrorMemoryAllocation;
  return tail::launch<4, 4, 2>(d, stream, status);
}

}  // namespace fn

extern "C" void fn_debug_set_tail_tc(int v) { fn::tail::g_enable = v != 0; }
