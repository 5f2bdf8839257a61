// tcgen05 / TMEM implicit-GEMM convolution for sm_100a.
//
// One CTA computes a tile of 16 x (8*J) output pixels for all N output channels:
//
//   gather     all 160 threads read the raw bf16 NHWC halo tile (18 x (8J+2) pixels), apply the
//              prologue nonlinearity (concat_elu & co, flow_architecture.py:14-29) in registers and
//              write the UMMA A operand to shared memory as 8-channel planes
//                   A[q][pv][pu][8]      q = k/8, 16 B per pixel, no swizzle
//              so that a 3x3 tap is a pure start-address offset of the shared-memory descriptor:
//              row r = 8*v+u of the M=128 MMA reads pixel (v+dv, u+du) -> byte (v*PU + u)*16 +
//              (dv*PU+du)*16, i.e. SBO = PU*16, LBO = plane stride.  The 2C-channel concat_elu
//              tensor and im2col never exist in HBM.
//   weights    pre-packed [kstep][2][N][8] bf16 (fn_pack_conv_weights_tc); streamed by the TMA
//              engine (cp.async.bulk + mbarrier) through a ring of <= 3 chunks of <= 32 KB
//   mma        one thread issues tcgen05.mma (M=128, N, K=16) per (k-step, sub-tile); fp32
//              accumulators live in TMEM, J*N columns; the nin skip (flow_architecture.py:136-142)
//              is just extra k-steps on a second A operand with the centre tap
//   epilogue   4 warps read their TMEM quadrant (tcgen05.ld), add bias, apply gate*sigmoid +
//              shortcut (avg-pool / front channel pad, :150-165) and store bf16 NHWC, 16 B/lane
//
// Reference semantics: tf.nn.conv2d 'SAME' stride 1 (flow_architecture.py:57-68) inside res_block
// (:129-165).  Stride-2 and transposed convs are in conv_tc_strided.cu (same machinery).
#include "common.cuh"
#include "ptx.cuh"

namespace fn {

constexpr int TC_THREADS = 160;      // 4 gather/epilogue warps + 1 MMA/TMA warp (which also gathers)
constexpr int TC_MAX_STAGES = 3;
constexpr int TC_CHUNK_BYTES = 32768;
constexpr int TC_SMEM_HEADER = 128;  // mbarriers + TMEM base pointer

// debug knobs (tests only): swap the LBO/SBO descriptor fields
static int g_tc_swap_lbo_sbo = 0;
static int* g_tc_status_dev = nullptr;   // device word: 0 ok, else first failing wait code

struct TcParams {
  const __nv_bfloat16* x;
  const __nv_bfloat16* skip;
  const __nv_bfloat16* wpk;
  const float* bias;
  void* y;
  const __nv_bfloat16* res;
  int* status;
  int B, H, W, C, Cs, SH, SW;
  int pro, epi;
  int N, Cout, F;
  int Cres, RH, RW, res_pool;
  int transposed, J;
  int EU, EV, tiles_u, tiles_v;
  int PU, P;
  int ntaps;            // 9 or 1
  int ksteps_tap;       // Keff / 16
  int ksteps_skip;      // Kskip_eff / 16
  int ksteps_total;
  int spc, nchunks, nstages;
  int a_main_bytes, a_skip_bytes;
  int tmem_cols;
  int swap_desc;
};

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, int swap) {
  return swap ? ptx::smem_desc(addr, sbo, lbo) : ptx::smem_desc(addr, lbo, sbo);
}

// gather one operand (main input or skip) into A planes
__device__ __forceinline__ void gather_operand(const TcParams& p, const __nv_bfloat16* __restrict__ src, int Craw,
                                               int SHh, int SWw, int b, int u0, int v0, unsigned char* planes) {
  const int chunks = Craw >> 3;
  const int items = chunks * p.P;
  const int mult = pro_mult(p.pro);
  for (int idx = threadIdx.x; idx < items; idx += TC_THREADS) {
    const int ch = idx / p.P;
    const int pp = idx - ch * p.P;
    const int pv = pp / p.PU;
    const int pu = pp - pv * p.PU;
    const int cu = u0 + pu - 1, cv = v0 + pv - 1;
    const int yy = p.transposed ? cu : cv;
    const int xx = p.transposed ? cv : cu;
    bf16x8 raw;
    raw.v = make_uint4(0u, 0u, 0u, 0u);
    if (cu >= 0 && cv >= 0 && yy < SHh && xx < SWw) {
      raw.v = __ldg(reinterpret_cast<const uint4*>(src + (((size_t)b * SHh + yy) * SWw + xx) * Craw + ch * 8));
    }
    bf16x8 o0, o1;
    if (p.pro == FN_PRO_NONE) {
      o0.v = raw.v;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float a0, a1;
        pro_apply(p.pro, __bfloat162float(raw.h[j]), a0, a1);
        o0.h[j] = __float2bfloat16(a0);
        o1.h[j] = __float2bfloat16(a1);
      }
    }
    *reinterpret_cast<uint4*>(planes + ((size_t)ch * p.P + pp) * 16) = o0.v;
    if (mult == 2) *reinterpret_cast<uint4*>(planes + ((size_t)(chunks + ch) * p.P + pp) * 16) = o1.v;
  }
}

// shortcut(res)[pixel, c0..c0+8) added into v[] (avg-pool and front channel pad, :153-165)
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

__global__ void __launch_bounds__(TC_THREADS) k_conv_tc(const __grid_constant__ TcParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  // header: [0]=done, [1..3]=full[s], [4..6]=empty[s] (8 B each), [7*8]=tmem base (4 B)
  const uint32_t smem_base = ptx::smem_u32(smem);
  const uint32_t bar_done = smem_base;
  const uint32_t bar_full = smem_base + 8;
  const uint32_t bar_empty = smem_base + 8 + 8 * TC_MAX_STAGES;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + 64);
  unsigned char* a_main = smem + TC_SMEM_HEADER;
  unsigned char* a_skip = a_main + p.a_main_bytes;
  unsigned char* b_ring = a_skip + p.a_skip_bytes;
  const uint32_t b_ring_u32 = ptx::smem_u32(b_ring);
  const uint32_t chunk_bytes = (uint32_t)p.spc * p.N * 32u;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

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
      ptx::mbar_init(bar_done, 1);
      for (int s = 0; s < TC_MAX_STAGES; ++s) {
        ptx::mbar_init(bar_full + 8 * s, 1);
        ptx::mbar_init(bar_empty + 8 * s, 1);
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

  // ---- gather + prologue: raw NHWC -> A planes
  gather_operand(p, p.x, p.C, p.H, p.W, b, u0, v0, a_main);
  if (p.skip != nullptr) gather_operand(p, p.skip, p.Cs, p.SH, p.SW, b, u0, v0, a_skip);

  ptx::fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the tensor-core (async) proxy
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  bool ok = true;
  if (warp == 4) {
    // ================= MMA issuer + weight-ring producer (one thread) =================
    if (lane == 0) {
      const uint32_t idesc = ptx::idesc_bf16_f32(128, p.N);
      const uint32_t a_lbo = (uint32_t)p.P * 16u, a_sbo = (uint32_t)p.PU * 16u;
      const uint32_t b_lbo = (uint32_t)p.N * 16u, b_sbo = 128u;
      const uint32_t a_main_u32 = ptx::smem_u32(a_main), a_skip_u32 = ptx::smem_u32(a_skip);
      const uint32_t centre = (uint32_t)(p.PU + 1) * 16u;
      const int main_steps = p.ntaps * p.ksteps_tap;
      for (int c = 0; c < p.nchunks && ok; ++c) {
        const int s = c % p.nstages;
        ok = ptx::mbar_wait(bar_full + 8 * s, (uint32_t)((c / p.nstages) & 1));
        if (!ok) { atomicCAS(p.status, 0, 2); break; }
        ptx::tc_fence_after_sync();
        const int steps = min(p.spc, p.ksteps_total - c * p.spc);
        for (int i = 0; i < steps; ++i) {
          const int ks = c * p.spc + i;
          uint32_t a_addr;
          if (ks < main_steps) {
            const int tap = ks / p.ksteps_tap, kk = ks - tap * p.ksteps_tap;
            uint32_t tapoff = centre;
            if (p.ntaps == 9) {
              const int di = tap / 3, dj = tap - 3 * di;
              const int dv = p.transposed ? dj : di, du = p.transposed ? di : dj;
              tapoff = (uint32_t)(dv * p.PU + du) * 16u;
            }
            a_addr = a_main_u32 + (uint32_t)(2 * kk) * a_lbo + tapoff;
          } else {
            const int kk = ks - main_steps;
            a_addr = a_skip_u32 + (uint32_t)(2 * kk) * a_lbo + centre;
          }
          const uint64_t bdesc = make_desc(b_ring_u32 + s * chunk_bytes + (uint32_t)i * p.N * 32u, b_lbo, b_sbo, p.swap_desc);
          for (int j = 0; j < p.J; ++j) {
            const uint64_t adesc = make_desc(a_addr + (uint32_t)j * 128u, a_lbo, a_sbo, p.swap_desc);
            ptx::mma_bf16_ss(tmem_base + (uint32_t)(j * p.N), adesc, bdesc, idesc, ks > 0 ? 1u : 0u);
          }
        }
        ptx::mma_commit(bar_empty + 8 * s);   // frees this ring slot once the MMAs above have read it
        // refill the slot of the PREVIOUS chunk (its MMAs are ahead of ours in the pipe) with chunk c-1+nstages
        const int cp = c - 1, cn = cp + p.nstages;
        if (cp >= 0 && cn < p.nchunks) {
          const int sp = cp % p.nstages;
          ok = ptx::mbar_wait(bar_empty + 8 * sp, (uint32_t)((cp / p.nstages) & 1));
          if (!ok) { atomicCAS(p.status, 0, 1); break; }
          const int nsteps = min(p.spc, p.ksteps_total - cn * p.spc);
          const uint32_t bytes = (uint32_t)nsteps * p.N * 32u;
          ptx::mbar_arrive_expect_tx(bar_full + 8 * sp, bytes);
          ptx::bulk_g2s(b_ring_u32 + sp * chunk_bytes,
                        reinterpret_cast<const unsigned char*>(p.wpk) + (size_t)cn * chunk_bytes, bytes, bar_full + 8 * sp);
        }
      }
      ptx::mma_commit(bar_done);              // all MMAs complete -> accumulators readable
    }
    __syncwarp();
  } else {
    // ================= epilogue: TMEM -> registers -> bf16 NHWC =================
    ok = ptx::mbar_wait(bar_done, 0);
    if (!ok) atomicCAS(p.status, 0, 3);
    ptx::tc_fence_after_sync();
    const int r = threadIdx.x;            // accumulator row == TMEM lane
    const int v = r >> 3, u = r & 7;
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int j = 0; j < p.J; ++j) {
      const int cu = u0 + 8 * j + u, cv = v0 + v;
      const bool live = ok && cu < p.EU && cv < p.EV;
      const int oy = p.transposed ? cu : cv, ox = p.transposed ? cv : cu;
      const size_t pix = ((size_t)b * p.H + oy) * p.W + ox;
      const uint32_t col0 = lane_base + (uint32_t)(j * p.N);
      if (p.epi == FN_EPI_GATE_RES) {
        for (int c0 = 0; c0 < p.F; c0 += 8) {
          float a[8], g[8];
          ptx::tmem_ld8x2(col0 + c0, col0 + p.F + c0, a, g);
          float o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            o[i] = (a[i] + __ldg(p.bias + c0 + i)) * sigmoid_f(g[i] + __ldg(p.bias + p.F + c0 + i));
          if (live) {
            add_shortcut8(p, b, oy, ox, c0, o);
            bf16x8 st;
#pragma unroll
            for (int i = 0; i < 8; ++i) st.h[i] = __float2bfloat16(o[i]);
            *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.y) + pix * p.F + c0) = st.v;
          }
        }
      } else {
        for (int c0 = 0; c0 < p.F; c0 += 8) {
          float o[8];
          ptx::tmem_ld8(col0 + c0, o);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] += __ldg(p.bias + c0 + i);
          if (live) {
            if (p.epi == FN_EPI_RES) add_shortcut8(p, b, oy, ox, c0, o);
            bf16x8 st;
#pragma unroll
            for (int i = 0; i < 8; ++i) st.h[i] = __float2bfloat16(o[i]);
            *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.y) + pix * p.F + c0) = st.v;
          }
        }
      }
    }
  }

  // ---- teardown
  ptx::tc_fence_before_sync();
  __syncthreads();
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

static bool plan_conv_tc(const fn_conv_desc& d, TcParams& p, size_t& smem_bytes, const char** why) {
  static const char* w_stride = "stride != 1 (handled by the strided kernel)";
  static const char* w_k = "K per tap not a multiple of 16";
  static const char* w_epi = "epilogue has no tcgen05 variant";
  static const char* w_drop = "dropout prologue";
  static const char* w_n = "N out of range";
  static const char* w_smem = "tile does not fit shared memory";
  const char* dummy;
  if (!why) why = &dummy;
  if (d.stride != 1) { *why = w_stride; return false; }
  if (d.epilogue == FN_EPI_TANH) { *why = w_epi; return false; }
  if (d.keep_prob < 1.0f) { *why = w_drop; return false; }
  if (d.w_tc == nullptr) { *why = "w_tc is null"; return false; }
  const int mult = pro_mult(d.prologue);
  const int Keff = d.Cin * mult;
  const int Kskip = d.skip ? d.Cskip * mult : 0;
  if ((d.Cin % 8) != 0 || (Keff % 16) != 0 || (Kskip % 16) != 0) { *why = w_k; return false; }
  const int N = round_up(d.Cout, 16);
  if (N < 16 || N > 256) { *why = w_n; return false; }
  const bool gated = d.epilogue == FN_EPI_GATE_RES;

  memset(&p, 0, sizeof(p));
  p.x = (const __nv_bfloat16*)d.x;
  p.skip = (const __nv_bfloat16*)d.skip;
  p.wpk = (const __nv_bfloat16*)d.w_tc;
  p.bias = d.bias;
  p.y = d.y;
  p.res = (const __nv_bfloat16*)d.res;
  p.B = d.B; p.H = d.H; p.W = d.W; p.C = d.Cin; p.Cs = d.Cskip; p.SH = d.SH; p.SW = d.SW;
  p.pro = d.prologue; p.epi = d.epilogue;
  p.N = N; p.Cout = d.Cout; p.F = gated ? d.Cout / 2 : d.Cout;
  p.Cres = d.Cres; p.RH = d.RH; p.RW = d.RW; p.res_pool = d.res_pool;
  p.ntaps = d.ksize * d.ksize;
  p.ksteps_tap = Keff / 16;
  p.ksteps_skip = Kskip / 16;
  p.ksteps_total = p.ntaps * p.ksteps_tap + p.ksteps_skip;
  p.spc = TC_CHUNK_BYTES / (N * 32);
  if (p.spc < 1) p.spc = 1;
  if (p.spc > p.ksteps_total) p.spc = p.ksteps_total;
  p.nchunks = (p.ksteps_total + p.spc - 1) / p.spc;
  p.nstages = p.nchunks < TC_MAX_STAGES ? p.nchunks : TC_MAX_STAGES;
  p.swap_desc = g_tc_swap_lbo_sbo;

  // orientation: the 8-wide fast axis u runs along x unless the image is too short for 16 rows
  p.transposed = (d.H < 16 && d.W >= 16) ? 1 : 0;
  p.EU = p.transposed ? d.H : d.W;
  p.EV = p.transposed ? d.W : d.H;
  p.tiles_v = (p.EV + 15) / 16;

  const int keff_planes = Keff / 8, kskip_planes = Kskip / 8;
  const long long want_ctas = 2LL * 148;
  int bestJ = 0;
  for (int J = 4; J >= 1; J >>= 1) {   // largest J that still gives two waves of CTAs, else the smallest feasible
    if (J * N > 256) continue;
    if (J > 1 && (p.EU % (8 * J)) != 0) continue;
    const int PU = 8 * J + 2, P = 18 * PU;
    const size_t sm = TC_SMEM_HEADER + (size_t)(keff_planes + kskip_planes) * P * 16 + (size_t)p.nstages * p.spc * N * 32;
    if (sm > 200 * 1024) continue;
    bestJ = J;
    const long long ctas = (long long)d.B * p.tiles_v * ((p.EU + 8 * J - 1) / (8 * J));
    if (ctas >= want_ctas) break;
  }
  if (bestJ == 0) { *why = w_smem; return false; }
  p.J = bestJ;
  p.PU = 8 * p.J + 2;
  p.P = 18 * p.PU;
  p.tiles_u = (p.EU + 8 * p.J - 1) / (8 * p.J);
  p.a_main_bytes = keff_planes * p.P * 16;
  p.a_skip_bytes = kskip_planes * p.P * 16;
  p.tmem_cols = pow2_cols(p.J * N);
  smem_bytes = TC_SMEM_HEADER + (size_t)p.a_main_bytes + p.a_skip_bytes + (size_t)p.nstages * p.spc * N * 32;
  return true;
}

bool conv_tc_supported(const fn_conv_desc& d, const char** why) {
  TcParams p;
  size_t sm;
  return plan_conv_tc(d, p, sm, why);
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

int conv_fwd_tc(const fn_conv_desc& d, cudaStream_t stream) {
  TcParams p;
  size_t smem_bytes = 0;
  const char* why = "";
  if (!plan_conv_tc(d, p, smem_bytes, &why)) {
    set_error("conv_fwd_tc: unsupported: %s", why);
    return FN_E_UNSUPPORTED;
  }
  int rc = ensure_status_word();
  if (rc) return rc;
  p.status = g_tc_status_dev;
  static size_t configured = 0;
  if (smem_bytes > configured) {
    cudaError_t e = cudaFuncSetAttribute(k_conv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024));
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = 220 * 1024;
  }
  const unsigned grid = (unsigned)((long long)d.B * p.tiles_u * p.tiles_v);
  k_conv_tc<<<grid, TC_THREADS, smem_bytes, stream>>>(p);
  count_launch();
  return check_launch("k_conv_tc");
}

// ---- weight packing: src [T][Keff][Cout] (+ [Kskip][Cout]) bf16 -> dst [kstep][2][N][8] bf16
__global__ void k_pack_tc(const __nv_bfloat16* __restrict__ w, const __nv_bfloat16* __restrict__ ws,
                          __nv_bfloat16* __restrict__ dst, int T, int Keff, int Kskip, int Cout, int N) {
  const int main_steps = T * (Keff / 16);
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
        const int tap = ks / (Keff / 16), kk = ks % (Keff / 16);
        const int k = kk * 16 + h * 8 + e;
        v = w[((size_t)tap * Keff + k) * Cout + n];
      } else {
        const int k = (ks - main_steps) * 16 + h * 8 + e;
        v = ws[(size_t)k * Cout + n];
      }
    }
    dst[i] = v;
  }
}

long long conv_weights_tc_bytes(int ksize, int Keff, int Kskip, int Cout, int gated) {
  (void)gated;
  if ((Keff % 16) != 0 || (Kskip % 16) != 0 || Cout <= 0) return 0;
  const long long steps = (long long)ksize * ksize * (Keff / 16) + Kskip / 16;
  return steps * round_up(Cout, 16) * 32;
}

int pack_conv_weights_tc(const void* w, const void* w_skip, void* dst, int ksize, int Keff, int Kskip, int Cout,
                         int gated, cudaStream_t stream) {
  (void)gated;
  FN_REQUIRE((Keff % 16) == 0 && (Kskip % 16) == 0, FN_E_ALIGN, "pack_conv_weights_tc: K (%d,%d) must be multiples of 16",
             Keff, Kskip);
  FN_REQUIRE(Kskip == 0 || w_skip != nullptr, FN_E_INVAL, "pack_conv_weights_tc: null w_skip");
  const int N = round_up(Cout, 16);
  const long long total = conv_weights_tc_bytes(ksize, Keff, Kskip, Cout, gated) / 2;
  const int grid = (int)((total + 255) / 256);
  k_pack_tc<<<grid, 256, 0, stream>>>((const __nv_bfloat16*)w, (const __nv_bfloat16*)w_skip, (__nv_bfloat16*)dst,
                                      ksize * ksize, Keff, Kskip, Cout, N);
  count_launch();
  return check_launch("k_pack_tc");
}

// transposed / strided tensor-core variants are not built yet: the CUDA-core kernels run them
bool tconv_tc_supported(const fn_tconv_desc&, const char** why) {
  if (why) *why = "transposed conv has no tcgen05 kernel yet";
  return false;
}
int tconv_fwd_tc(const fn_tconv_desc&, cudaStream_t) {
  set_error("tconv_fwd_tc: not built");
  return FN_E_UNSUPPORTED;
}

}  // namespace fn

// ---- debug / test hooks (not part of the drop-in surface, but exported for tests)
extern "C" {
void fn_debug_set_swap_lbo_sbo(int v) { fn::g_tc_swap_lbo_sbo = v; }
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
