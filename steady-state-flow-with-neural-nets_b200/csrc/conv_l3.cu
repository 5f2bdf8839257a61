// The stride-1 part of the 32x64 level of flow_net in TWO launches, the organisation of conv_l4.cu one level up: a
// thread-block cluster of two CTAs per image, the image split by COLUMNS (32 x 32 pixels = EIGHT M = 128 UMMA tiles per
// CTA, TMEM exactly full with N = 64 accumulators), activations resident in shared memory between the convolutions, the
// one-pixel halo column of every intermediate exchanged through distributed shared memory.
//
//   MODE_DOWN  x1 = conv_1 output of "resnet_3_downsample" [B,32,64,32] (the stride-2 conv stays a launch of its own: the
//              parity sub-planes of its 64x128 input do not fit), x2 [B,64,128,16] for the pooled shortcut
//              -> gated conv_2 (+ pooled / front-padded shortcut) -> res_block "resnet_3_0" -> x3 [B,32,64,32]
//              (reference flow_architecture.py:198-201)
//   MODE_UP    x1 = conv_1 (+ nin) output of "resnet_up_2_0" [B,32,64,32], z = up_conv_2's output (the block's input)
//              -> gated conv_2 (+ z) -> transposed conv "up_conv_3" -> [B,64,128,16]      (reference :223-228)
//
// Why: on this level a launch is 1024 tiles for 148 SMs -- 7 tiles per CTA, 12 k clocks of tcgen05.mma behind ~8 us of
// launch, fill and drain; layer by layer the five convolutions cost 105 us for 31 us of MMA time.
//
//   warp 16     producer   the stage weight images (37-74 KB each) through a 4 x 16 KB ring
//   warps 0-15  workers    warps 0 / 1 each issue the MMAs of one row of four tiles (the four share every B k-step);
//                          all sixteen run the epilogues: warp w owns TMEM lane quadrant w & 3 of the two tiles in tile
//                          column w >> 2
//
// Planes: [plane = 8 channels][row + 1][col + 1][8], 34 x 34 pixels per plane, zero border (TF 'SAME'); UMMA row
// r = 8 row + (col & 7) inside a 16 x 8 tile.  Same operands, bf16 rounding points and MMA order per accumulator as the
// layer-by-layer kernels.
#include <cuda.h>

#include <type_traits>

#include "cluster.cuh"

namespace fn {
using namespace tma;
namespace l3 {
using namespace cl;

constexpr int NWORK = 16, WTHREADS = NWORK * 32, THREADS = WTHREADS + 32;
constexpr int NST = 4, CHUNK = 16384;
constexpr int HDR = 4096;                      // [0,128) mbarriers, [256] TMEM base, [512, ...) biases
constexpr int C2 = 16, C3 = 32, LH2 = 64, LW2 = 128, IH = 32, IW = 64, CW = 32;
constexpr int PU = 34, PV = 34, P = PU * PV, PPAD = 1161;        // 1161 = 1 (mod 8): conflict-free plane stride
constexpr int PLANES_BYTES = 8 * PPAD * 16;
constexpr int OFF_RING = HDR;
constexpr int OFF_A = OFF_RING + NST * CHUNK;
constexpr int SMEM = OFF_A + PLANES_BYTES;
static_assert(SMEM <= 231424, "shared memory");
// biases, floats
constexpr int B2 = 0, B3 = 64, B4 = 96, NBIAS_DOWN = 160;        // conv_2 down (u | v), conv_1, conv_2 (u | v)
constexpr int B6 = 0, B7 = 64, NBIAS_UP = 80;                    // conv_2 (u | v), up_conv_3
// weight streams, bytes per stage (k-steps x N x 32)
constexpr int W2 = 36 * 64 * 32, W3 = 36 * 32 * 32, W4 = W2, W6 = W2, W7 = 18 * 16 * 32;
__host__ __device__ constexpr int chunks_of(int bytes) { return (bytes + CHUNK - 1) / CHUNK; }
constexpr int G2 = 0, G3 = G2 + chunks_of(W2), G4 = G3 + chunks_of(W3);
constexpr int G6 = 0, G7 = G6 + chunks_of(W6);

enum { SB_FULL = 0, SB_EMPTY = NST, SB_TF = 2 * NST, SB_XR = 2 * NST + 1, SB_XW = 2 * NST + 2 };
enum { MODE_DOWN = 0, MODE_UP = 1 };

struct Params {
  const __nv_bfloat16* x;       // x1 [B,32,64,32]: the block's conv_1 output
  const __nv_bfloat16* aux;     // DOWN: x2 [B,64,128,16] (pooled shortcut); UP: z [B,32,64,32] (shortcut)
  const unsigned char* wstream; // the stage weight images back to back
  const float* bias;
  __nv_bfloat16* y;             // DOWN: x3 [B,32,64,32]; UP: [B,64,128,16]
  int* status;
  long long* timing;            // debug: 16 clock64 stamps per CTA (fn_debug_set_timing; nullptr in production)
  int timing_ctas;
  int B;
};

__device__ __forceinline__ void worker_barrier() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1) k_level3(const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = ptx::smem_u32(smem);
  auto bar = [&](int i) { return sbase + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + 256);
  float* sbias = reinterpret_cast<float*>(smem + 512);
  const uint32_t sb = sbase + 512u;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int tid = threadIdx.x;
  const int b = (int)blockIdx.x >> 1;
  const uint32_t rank = cluster_rank(), peer = rank ^ 1u;
  constexpr int NBIAS = MODE == MODE_DOWN ? NBIAS_DOWN : NBIAS_UP;

  if (warp == NWORK) {
    ptx::tmem_alloc(sbase + 256, 512u);
    ptx::tmem_relinquish();
    if (lane == 0) {
      for (int s = 0; s < NST; ++s) {
        ptx::mbar_init(bar(SB_FULL + s), 1);
        ptx::mbar_init(bar(SB_EMPTY + s), 2);          // one commit per issuing warp
      }
      ptx::mbar_init(bar(SB_TF), 2);
      ptx::mbar_init(bar(SB_XR), 1);
      ptx::mbar_init(bar(SB_XW), NWORK);
      ptx::fence_mbar_init();
    }
    __syncwarp();
  }
  for (int i = tid; i < NBIAS; i += THREADS) {
    float bv = __ldg(p.bias + i);
    // the v halves of the gated convs: sigmoid(z) = 0.5 + 0.5 tanh(z / 2)
    const bool vhalf = MODE == MODE_DOWN ? ((i >= B2 + 32 && i < B3) || i >= B4 + 32) : (i >= B6 + 32 && i < B7);
    if (vhalf) bv *= 0.5f;
    sbias[i] = bv;
  }
  ptx::tc_fence_before_sync();
  cluster_sync_all();                 // both CTAs' barriers exist before anything arrives on them
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  bool ok = true;
  long long* stamps = (p.timing != nullptr && (int)blockIdx.x < p.timing_ctas && tid == 0) ? p.timing + (size_t)blockIdx.x * 16 : nullptr;
  int nstamp = 0;
  auto stamp = [&]() { if (stamps) stamps[nstamp++] = clock64(); };
  stamp();

  if (warp == NWORK) {
    // =========================== producer: the weight stream (both CTAs read all of it) ===========================
    if (lane == 0) {
      constexpr int NSTAGE = MODE == MODE_DOWN ? 3 : 2;
      const int stage_bytes[3] = {MODE == MODE_DOWN ? W2 : W6, MODE == MODE_DOWN ? W3 : W7, W4};
      const uint64_t pol = ptx::policy_evict_last();
      int gc = 0;
      size_t off = 0;
      for (int st = 0; st < NSTAGE && ok; ++st) {
        for (int rem = stage_bytes[st]; rem > 0 && ok; ++gc) {
          const int bytes = rem < CHUNK ? rem : CHUNK;
          const int s = gc % NST;
          if (gc >= NST) {
            ok = ptx::mbar_wait(bar(SB_EMPTY + s), (uint32_t)(((gc / NST) - 1) & 1));
            if (!ok) { atomicCAS(p.status, 0, 91); break; }
          }
          ptx::mbar_arrive_expect_tx(bar(SB_FULL + s), (uint32_t)bytes);
          ptx::bulk_g2s_hint(sbase + OFF_RING + s * CHUNK, p.wstream + off, (uint32_t)bytes, bar(SB_FULL + s), pol);
          off += (size_t)bytes;
          rem -= bytes;
        }
      }
    }
    __syncwarp();
  } else {
    // =========================== workers ===========================
    const uint32_t leader = ptx::elect_one();
    const int q = warp & 3, tc = warp >> 2;           // TMEM lane quadrant, tile column
    const int r = q * 32 + lane;                      // accumulator row == TMEM lane
    const int col = tc * 8 + (r & 7);                 // this thread's pixels: (tr * 16 + (r >> 3), col), tr = 0, 1
    const int gcol = (int)rank * CW + col;            // column inside the image
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t A = sbase + OFF_A;
    const uint32_t A_peer = map_to_cta(A, peer);
    const uint32_t xr_peer = map_to_cta(bar(SB_XR), peer), xw_peer = map_to_cta(bar(SB_XW), peer);
    auto pp16_of = [&](int tr) { return (uint32_t)((tr * 16 + (r >> 3) + 1) * PU + (col + 1)) * 16u; };
    // the CTA's border column is the peer's halo: rank 0's column 31 -> the peer's plane column 0, rank 1's column 0 -> 33
    const bool border = rank == 0 ? col == CW - 1 : col == 0;
    auto pp16_peer_of = [&](int tr) { return (uint32_t)((tr * 16 + (r >> 3) + 1) * PU + (rank == 0 ? 0 : PU - 1)) * 16u; };
    uint32_t tf_par = 0, x_par = 0;

    auto chunk_ready = [&](int gc, uint32_t lbo) -> uint64_t {
      const int s = gc % NST;
      ok = ok && ptx::mbar_wait(bar(SB_FULL + s), (uint32_t)((gc / NST) & 1));
      if (!ok) atomicCAS(p.status, 0, 92);
      ptx::tc_fence_after_sync();
      return ptx::smem_desc(sbase + OFF_RING + s * CHUNK, lbo, 128);
    };
    // everybody: this stage's accumulators are complete; the peer learns that this CTA's planes may be overwritten.
    // readers_first: other warps read the operand planes during this stage's MMAs (the shortcut recovered from them), so
    // nobody -- neither this CTA's epilogue nor the peer -- may overwrite them before every warp has passed this point
    auto stage_done = [&](bool readers_first = false) {
      ok = ok && ptx::mbar_wait(bar(SB_TF), tf_par);
      if (!ok) atomicCAS(p.status, 0, 93);
      tf_par ^= 1u;
      ptx::tc_fence_after_sync();
      if (readers_first) worker_barrier();
      if (tid == 0) mbar_arrive_remote(xr_peer);
    };
    auto peer_planes_free = [&]() {
      ok = ok && mbar_wait_cluster(bar(SB_XR), x_par);
      if (!ok) atomicCAS(p.status, 0, 94);
    };
    auto planes_exchanged = [&]() {
      fence_proxy_async_all();
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(xw_peer);
      worker_barrier();
      ok = ok && mbar_wait_cluster(bar(SB_XW), x_par);
      if (!ok) atomicCAS(p.status, 0, 95);
      fence_proxy_async_all();            // the peer's generic-proxy writes, acquired above, before this CTA's tensor-core reads
      x_par ^= 1u;
    };
    auto put_plane = [&](int tr, int plane, const uint4& v) {
      sts128(A + (uint32_t)(plane * PPAD * 16) + pp16_of(tr), v);
      if (border) st_cluster128(A_peer + (uint32_t)(plane * PPAD * 16) + pp16_peer_of(tr), v);
    };

    // ---- MMAs of one 3x3 convolution on the planes, K = 64 per tap (4 k-steps): warp 0 / 1 issues tile row 0 / 1, four
    // tiles per B k-step; accumulator of tile t = 4 row + column at TMEM column t * N
    const uint64_t a_planes = ptx::smem_desc(A + (uint32_t)(warp * 16 * PU * 16), PPAD * 16, PU * 16);
    auto issue_conv = [&](auto nconst, int g0) {
      constexpr int N = decltype(nconst)::value;
      constexpr int SPC = CHUNK / (N * 32), TPC = SPC / 4;          // k-steps, taps per chunk
      constexpr int NCH = (9 + TPC - 1) / TPC;
      constexpr uint32_t idesc = ptx::idesc_bf16_f32(128, N);
      const uint32_t d_tmem = tmem_base + (uint32_t)(warp * 4 * N);
#pragma unroll 1
      for (int c = 0; c < NCH; ++c) {
        const uint64_t b_chunk = chunk_ready(g0 + c, N * 16);
#pragma unroll
        for (int t = 0; t < TPC; ++t) {
          const int tap = c * TPC + t;
          if (tap < 9) {
            const int di = tap / 3, dj = tap - 3 * di;
            const uint64_t a_tap = a_planes + (uint64_t)((uint32_t)((di * PU + dj) * 16) >> 4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (leader) ptx::mma_bf16_ss(d_tmem + (uint32_t)(j * N), a_tap + (uint64_t)((kk * 2u * PPAD * 16u + j * 128u) >> 4),
                                             b_chunk + (uint32_t)(((t * 4 + kk) * N * 32) >> 4), idesc, (tap == 0 && kk == 0) ? 0u : 1u);
              }
            }
          }
        }
        if (leader) ptx::mma_commit(bar(SB_EMPTY + (g0 + c) % NST));
      }
      if (leader) ptx::mma_commit(bar(SB_TF));
      __syncwarp();
    };
    using N32 = std::integral_constant<int, 32>;
    using N64 = std::integral_constant<int, 64>;

    // gate of 8 channels of tile t: (a + bu) * sigmoid(v + bv), product not yet rounded
    auto gate8 = [&](int boff, int t, int c, float (&a)[8], float (&sg)[8]) {
      float gt[8], bu[8], bv[8];
      lds_f8(sb + 4u * (uint32_t)(boff + c), bu);
      lds_f8(sb + 4u * (uint32_t)(boff + 32 + c), bv);
      ptx::tmem_ld8x2(lane_base + (uint32_t)(t * 64 + c), lane_base + (uint32_t)(t * 64 + 32 + c), a, gt);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a[i] += bu[i];
        sg[i] = fmaf(tanh_approx(fmaf(gt[i], 0.5f, bv[i])), 0.5f, 0.5f);
      }
    };

    // ---- first operand: x1 -> concat_elu -> planes, the halo column read directly, zero outside the image
    {
      const __nv_bfloat16* xb = p.x + (size_t)b * IH * IW * C3;
      constexpr int ITEMS = PV * PU * 4;
      constexpr int UN = (ITEMS + WTHREADS - 1) / WTHREADS;       // every load of a thread in flight at once
      uint4 in[UN];
#pragma unroll
      for (int k = 0; k < UN; ++k) {
        const int idx = tid + WTHREADS * k;
        const int ch = idx & 3, sp = idx >> 2;
        const int pr = sp / PU, pc = sp - pr * PU;
        const int h = pr - 1, w = (int)rank * CW + pc - 1;
        in[k] = make_uint4(0u, 0u, 0u, 0u);
        if (idx < ITEMS && h >= 0 && h < IH && w >= 0 && w < IW) in[k] = __ldg(reinterpret_cast<const uint4*>(xb + ((size_t)h * IW + w) * C3 + ch * 8));
      }
#pragma unroll
      for (int k = 0; k < UN; ++k) {
        const int idx = tid + WTHREADS * k;
        if (idx >= ITEMS) break;
        const int ch = idx & 3, sp = idx >> 2;
        uint4 o0, o1;
        concat_elu8(in[k], o0, o1);
        const uint32_t d0 = A + (uint32_t)(ch * PPAD + sp) * 16u;
        sts128(d0, o0);
        sts128(d0 + (uint32_t)(4 * PPAD * 16), o1);
      }
    }

    if constexpr (MODE == MODE_DOWN) {
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before_sync();
      worker_barrier();
      stamp();

      // ---- stage 2: conv_2 of the down block, gate, pooled shortcut in the last 16 channels
      if (warp < 2) {
        ptx::tc_fence_after_sync();
        issue_conv(N64{}, G2);
      }
      // ---- the shortcut of this stage, front channel pad (reference :158-163): output channels 16.. take the 2x2 average pool
      // of x2.  Fetched while the stage's MMAs run: 16 scattered 16-byte loads per thread cost 8 k clocks of L1 tag time, which
      // the 15 k clocks of MMAs hide -- at this level the weight ring (74 KB per stage, prefetched four chunks deep) is not
      // the critical resource the same loads disturbed in conv_l4.cu
      float pool[2][2][8];
      {
        const __nv_bfloat16* x2b = p.aux + (size_t)b * LH2 * LW2 * C2;
#pragma unroll
        for (int tr = 0; tr < 2; ++tr) {
          const int row = tr * 16 + (r >> 3);
          const __nv_bfloat16* r00 = x2b + ((size_t)(2 * row) * LW2 + 2 * gcol) * C2;
          bf16x8 q4[2][4];
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            q4[g][0].v = __ldg(reinterpret_cast<const uint4*>(r00 + g * 8));
            q4[g][1].v = __ldg(reinterpret_cast<const uint4*>(r00 + g * 8 + C2));
            q4[g][2].v = __ldg(reinterpret_cast<const uint4*>(r00 + g * 8 + (size_t)LW2 * C2));
            q4[g][3].v = __ldg(reinterpret_cast<const uint4*>(r00 + g * 8 + (size_t)LW2 * C2 + C2));
          }
#pragma unroll
          for (int g = 0; g < 2; ++g)
#pragma unroll
            for (int i = 0; i < 8; ++i)
              pool[tr][g][i] = 0.25f * (__bfloat162float(q4[g][0].h[i]) + __bfloat162float(q4[g][1].h[i]) + __bfloat162float(q4[g][2].h[i]) +
                                        __bfloat162float(q4[g][3].h[i]));
        }
      }
      stage_done();
      stamp();
      peer_planes_free();
#pragma unroll
      for (int tr = 0; tr < 2; ++tr) {
        const int t = tr * 4 + tc;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float a[8], sg[8], o[8];
          gate8(B2, t, g * 8, a, sg);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = a[i] * sg[i];
          if (g >= 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] += pool[tr][g - 2][i];
          }
          uint4 o0, o1;
          concat_elu8(pack8(o), o0, o1);
          put_plane(tr, g, o0);
          put_plane(tr, 4 + g, o1);
        }
      }
      planes_exchanged();
      stamp();

      // ---- stage 3: conv_1 of resnet_3_0; its operand is [elu | elu(-.)] of the block input, from which the shortcut of
      // stage 4 is recovered while the MMAs read the same planes
      if (warp < 2) {
        ptx::tc_fence_after_sync();
        issue_conv(N32{}, G3);
      }
      uint4 res[2][4];
#pragma unroll
      for (int tr = 0; tr < 2; ++tr)
#pragma unroll
        for (int g = 0; g < 4; ++g)
          res[tr][g] = raw_from_concat_elu(lds128(A + (uint32_t)(g * PPAD * 16) + pp16_of(tr)), lds128(A + (uint32_t)((4 + g) * PPAD * 16) + pp16_of(tr)));
      stage_done(true);
      stamp();
      peer_planes_free();
#pragma unroll 1
      for (int tr = 0; tr < 2; ++tr) {
        const int t = tr * 4 + tc;
#pragma unroll
        for (int g = 0; g < 4; g += 2) {
          float o[2][8], bv[2][8];
          lds_f8(sb + 4u * (uint32_t)(B3 + g * 8), bv[0]);
          lds_f8(sb + 4u * (uint32_t)(B3 + g * 8 + 8), bv[1]);
          ptx::tmem_ld8x2(lane_base + (uint32_t)(t * 32 + g * 8), lane_base + (uint32_t)(t * 32 + g * 8 + 8), o[0], o[1]);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int i = 0; i < 8; ++i) o[h][i] += bv[h][i];
            uint4 o0, o1;
            concat_elu8(pack8(o[h]), o0, o1);
            put_plane(tr, g + h, o0);
            put_plane(tr, 4 + g + h, o1);
          }
        }
      }
      planes_exchanged();
      stamp();

      // ---- stage 4: conv_2 of resnet_3_0, gate, shortcut -> x3 in global memory
      if (warp < 2) {
        ptx::tc_fence_after_sync();
        issue_conv(N64{}, G4);
      }
      stage_done();
      stamp();
#pragma unroll
      for (int tr = 0; tr < 2; ++tr) {
        const int t = tr * 4 + tc;
        __nv_bfloat16* dst = p.y + (((size_t)b * IH + tr * 16 + (r >> 3)) * IW + gcol) * C3;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float a[8], sg[8], o[8];
          gate8(B4, t, g * 8, a, sg);
          bf16x8 rv;
          rv.v = res[tr][g];
          // product and shortcut add rounded separately, as in the layer-by-layer kernels
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = __fadd_rn(__fmul_rn(a[i], sg[i]), __bfloat162float(rv.h[i]));
          if (ok) *reinterpret_cast<uint4*>(dst + g * 8) = pack8(o);
        }
      }
    } else {
      // ---- the shortcut z of this thread's pixels, fetched with the operand
      uint4 zres[2][4];
#pragma unroll
      for (int tr = 0; tr < 2; ++tr) {
        const __nv_bfloat16* zp = p.aux + (((size_t)b * IH + tr * 16 + (r >> 3)) * IW + gcol) * C3;
#pragma unroll
        for (int g = 0; g < 4; ++g) zres[tr][g] = __ldg(reinterpret_cast<const uint4*>(zp + g * 8));
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before_sync();
      worker_barrier();
      stamp();

      // ---- stage 6: conv_2 of resnet_up_2_0, gate, shortcut z; the result feeds the transposed conv RAW (no
      // nonlinearity, reference :225): 4 planes
      if (warp < 2) {
        ptx::tc_fence_after_sync();
        issue_conv(N64{}, G6);
      }
      stage_done();
      stamp();
      peer_planes_free();
#pragma unroll
      for (int tr = 0; tr < 2; ++tr) {
        const int t = tr * 4 + tc;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float a[8], sg[8], o[8];
          gate8(B6, t, g * 8, a, sg);
          bf16x8 rv;
          rv.v = zres[tr][g];
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = __fadd_rn(__fmul_rn(a[i], sg[i]), __bfloat162float(rv.h[i]));
          put_plane(tr, g, pack8(o));
        }
      }
      planes_exchanged();
      stamp();

      // ---- stage 7: transposed conv as four output phases, K = 32 per tap (2 k-steps), N = 16 per phase;
      // all 18 k-steps are one ring chunk.  Accumulators: tile t, phase ph at columns t * 64 + ph * 16
      if (warp < 2) {
        ptx::tc_fence_after_sync();
        constexpr uint32_t idesc = ptx::idesc_bf16_f32(128, 16);
        const uint64_t b_chunk = chunk_ready(G7, 16 * 16);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int di = tap / 3, dj = tap - 3 * di;
          // out[2m + p] takes tap d = p from in[m] and, for p = 0, tap d = 2 from in[m - 1]
          const uint64_t a_tap = a_planes + (uint64_t)((uint32_t)((((di == 2) ? 0 : 1) * PU + ((dj == 2) ? 0 : 1)) * 16) >> 4);
          const uint32_t d_ph = (uint32_t)((((di == 1) ? 2 : 0) + ((dj == 1) ? 1 : 0)) * 16);
          const bool first_tap = tap == 0 || tap == 1 || tap == 3 || tap == 4;
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (leader) ptx::mma_bf16_ss(tmem_base + (uint32_t)((warp * 4 + j) * 64) + d_ph, a_tap + (uint64_t)((kk * 2u * PPAD * 16u + j * 128u) >> 4),
                                           b_chunk + (uint32_t)(((tap * 2 + kk) * 16 * 32) >> 4), idesc, (first_tap && kk == 0) ? 0u : 1u);
            }
          }
        }
        if (leader) ptx::mma_commit(bar(SB_EMPTY + G7 % NST));
        if (leader) ptx::mma_commit(bar(SB_TF));
        __syncwarp();
      }
      stage_done();
      stamp();
      {
        __nv_bfloat16* yb = p.y + (size_t)b * LH2 * LW2 * C2;
#pragma unroll 1
        for (int tr = 0; tr < 2; ++tr) {
          const int t = tr * 4 + tc, row = tr * 16 + (r >> 3);
#pragma unroll
          for (int ph = 0; ph < 4; ++ph) {
            const int oy = 2 * row + (ph >> 1), ox = 2 * gcol + (ph & 1);
            __nv_bfloat16* dst = yb + ((size_t)oy * LW2 + ox) * C2;
            float o[2][8], bv[2][8];
            lds_f8(sb + 4u * (uint32_t)B7, bv[0]);
            lds_f8(sb + 4u * (uint32_t)(B7 + 8), bv[1]);
            ptx::tmem_ld8x2(lane_base + (uint32_t)(t * 64 + ph * 16), lane_base + (uint32_t)(t * 64 + ph * 16 + 8), o[0], o[1]);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
              for (int i = 0; i < 8; ++i) o[h][i] += bv[h][i];
              if (ok) *reinterpret_cast<uint4*>(dst + 8 * h) = pack8(o[h]);
            }
          }
        }
      }
    }
  }

  stamp();
  // ---- neither CTA may leave while the other can still write into it or arrive on its barriers
  ptx::tc_fence_before_sync();
  cluster_sync_all();
  if (warp == NWORK) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, 512u);
  }
}

}  // namespace l3

int* tc_status_word();   // conv_tc.cu
void tc_timing_buffer(long long** buf, int* ctas);   // conv_tc.cu: debug stamp buffer (fn_debug_set_timing)

int level3_fwd(const fn_level3_desc& d, cudaStream_t stream) {
  using namespace l3;
  Params p;
  p.x = (const __nv_bfloat16*)d.x;
  p.aux = (const __nv_bfloat16*)d.aux;
  p.wstream = (const unsigned char*)d.w_stream;
  p.bias = d.bias;
  p.y = (__nv_bfloat16*)d.y;
  p.status = tc_status_word();
  if (p.status == nullptr) return (int)cudaErrorMemoryAllocation;
  p.B = d.B;
  tc_timing_buffer(&p.timing, &p.timing_ctas);
  p.timing_ctas /= 2;           // 16 stamps per CTA out of a buffer of 8 per registered CTA
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_level3<MODE_DOWN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_level3<MODE_UP>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(level3): %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  if (d.up) k_level3<MODE_UP><<<2 * d.B, THREADS, SMEM, stream>>>(p);
  else k_level3<MODE_DOWN><<<2 * d.B, THREADS, SMEM, stream>>>(p);
  count_launch();
  note_tcgen05(true);
  return check_launch("k_level3");
}

long long level3_weight_bytes(int up) {
  using namespace l3;
  return up ? (long long)W6 + W7 : (long long)W2 + W3 + W4;
}

}  // namespace fn
