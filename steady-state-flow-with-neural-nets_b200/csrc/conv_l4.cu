// The 16x32 level of flow_net in TWO launches around the fused 8x16-level kernel (conv_l5c.cu), a thread-block cluster of
// two CTAs per image, the image split by COLUMNS (16 x 16 pixels = two M = 128 UMMA tiles per CTA):
//
//   MODE_DOWN  x3 [B,32,64,32] -> res_block "resnet_4_downsample" (stride-2 conv_1, gated conv_2, pooled / front-padded
//              shortcut) -> res_block "resnet_4_0" -> x4 [B,16,32,64]                (reference flow_architecture.py:202-205)
//   MODE_UP    z = up_conv_1's output [B,16,32,64], skip a = x4 -> res_block "resnet_up_1_0" (conv_1 + nin(a), gated
//              conv_2) -> transposed conv "up_conv_2" -> [B,32,64,32]               (reference :217-222)
//
// Why: layer by layer these seven launches cost 122 us at batch 64 for 28 us of tcgen05.mma time -- every launch pays its
// start-up, a tensor-map round trip through HBM / L2 and a two-tile pipeline that never fills (profiles/r2_ncu_summary.json:
// tensor pipe 12-23 %, issue slots 12-17 %).  Here the activations of an image stay in shared memory between the
// convolutions, as in conv_l5.cu; what is new at this level is that an image is FOUR M = 128 tiles, too much for one CTA's
// shared memory next to a weight ring, so two CTAs share it and exchange the one-pixel halo column of every intermediate
// through distributed shared memory (st.shared::cluster), 4 KB per stage:
//
//   warp 16     producer   streams the stage's packed weight images (fn_pack_conv_weights_tc form, back to back) through a
//                          4 x 16 KB ring (cp.async.bulk + mbarrier), running ahead across stage boundaries
//   warps 0-15  workers    build the first operand from global memory, warp 0 elects one lane to issue the unrolled
//                          tcgen05.mma program (the two tiles of a CTA share every B k-step while it is in the ring),
//                          then all sixteen warps run the epilogue -- warp w owns TMEM lane quadrant w & 3 of tile
//                          (w >> 2) & 1 and channel half w >> 3 -- TMEM -> bias / gate / shortcut -> bf16 -> [elu | elu(-.)]
//                          planes of the NEXT stage (own shared memory, border column also into the peer's halo)
//
//   MODE_ALL   both halves AND the 8x16 level between them (the stages of conv_l5c_body.cuh run by warps 0-7 on their own
//              barrier set) in ONE launch: x3 -> [B,32,64,32].  x4 and z still pass through global memory (64 KB per image,
//              L2), written and re-read on either side of a cluster barrier; what the merge removes is two launches
//              (8-13 us each inside the forward: 217 KB of shared memory per CTA means the next kernel starts only when
//              the previous one has drained) -- levels 4 + 5 = 12 convolutions in one launch
//
// Planes: [plane = 8 channels][row + 1][col + 1][8] with a zero border (TF 'SAME'), 18 x 18 pixels per plane; UMMA row
// r = 8 row + (col & 7): a tap is a start-address offset of the shared-memory matrix descriptor, exactly as in the
// layer-by-layer kernels, whose operands, bf16 rounding points and MMA order per accumulator this kernel reproduces.
#include <cuda.h>

#include <type_traits>

#include "common.cuh"
#include "conv_l5c_body.cuh"
#include "ptx.cuh"
#include "tma.cuh"

namespace fn {
using namespace tma;
namespace l4 {

constexpr int NWORK = 16, WTHREADS = NWORK * 32, THREADS = WTHREADS + 32;
constexpr int NST = 4, CHUNK = 16384;
constexpr int HDR = 4096;                      // [0,128) mbarriers, [256] TMEM base, [512, 512 + 384*4) biases
constexpr int C3 = 32, C4 = 64, LH3 = 32, LW3 = 64, IH = 16, IW = 32, CW = 16;
// MODE_DOWN stage 1 operand: the four (row, column)-parity sub-planes of concat_elu(x3) this CTA's 16 output columns
// read, each [8 planes][17 rows][17 cols][8]; 289 = 1 (mod 8): conflict-free plane stride
constexpr int PS1 = 17, P1 = PS1 * PS1;
constexpr int SUB1_BYTES = 8 * P1 * 16;
// stride-1 operand: [16 planes][18 rows][18 cols][8], plane stride padded to 1 (mod 8)
constexpr int PU = 18, PV = 18, P = PU * PV, PPAD = 329;
constexpr int PLANES_BYTES = 16 * PPAD * 16;
// MODE_UP nin operand: concat_elu(a), centre pixels only
constexpr int PS = 256, PSPAD = 257;
constexpr int SKIP_BYTES = 16 * PSPAD * 16;
constexpr int OFF_RING = HDR;
constexpr int OFF_A = OFF_RING + NST * CHUNK;
constexpr int OFF_R1 = OFF_A + PLANES_BYTES;   // MODE_DOWN: raw output of resnet_4_downsample, [256 px][64 ch], 16-byte units swizzled
constexpr int OFF_S = OFF_A + PLANES_BYTES;    // MODE_UP: the nin operand
constexpr int SMEM_DOWN = OFF_A + 4 * SUB1_BYTES;
constexpr int SMEM_UP = OFF_S + SKIP_BYTES;
static_assert(OFF_R1 + 256 * 128 <= SMEM_DOWN, "shortcut copy must fit behind the planes");
static_assert(SMEM_DOWN <= 232448 && SMEM_UP <= 232448, "shared memory");
// biases, floats
constexpr int B1 = 0, B2 = 64, B3 = 192, B4 = 256, NBIAS_DOWN = 384;       // conv_1 down, conv_2 down (u | v), conv_1, conv_2
constexpr int B5 = 0, B6 = 64, B7 = 192, NBIAS_UP = 224;                   // conv_1 + nin, conv_2 (u | v), up_conv_2
// weight streams, bytes per stage (k-steps x N x 32)
constexpr int W1 = 36 * 64 * 32, W2 = 72 * 128 * 32, W3 = 72 * 64 * 32, W4 = W2;
constexpr int W5 = 80 * 64 * 32, W6 = W2, W7 = 36 * 32 * 32;
__host__ __device__ constexpr int chunks_of(int bytes) { return (bytes + CHUNK - 1) / CHUNK; }
constexpr int G1 = 0, G2 = G1 + chunks_of(W1), G3 = G2 + chunks_of(W2), G4 = G3 + chunks_of(W3);
constexpr int G5 = 0, G6 = G5 + chunks_of(W5), G7 = G6 + chunks_of(W6);

enum { SB_FULL = 0, SB_EMPTY = NST, SB_TF = 2 * NST, SB_XR = 2 * NST + 1, SB_XW = 2 * NST + 2 };
enum { MODE_DOWN = 0, MODE_UP = 1, MODE_ALL = 2 };
constexpr int SMEM_ALL = l5c::SMEM > SMEM_DOWN ? (l5c::SMEM > SMEM_UP ? l5c::SMEM : SMEM_UP) : (SMEM_DOWN > SMEM_UP ? SMEM_DOWN : SMEM_UP);
static_assert(SMEM_ALL <= 231424, "shared memory");
constexpr int NCHUNK_DOWN = chunks_of(W1) + chunks_of(W2) + chunks_of(W3) + chunks_of(W4);
constexpr int BAR5 = 16;                       // MODE_ALL: first mbarrier of the level-5 stages' own set

struct Params {
  const __nv_bfloat16* x;       // DOWN / ALL: x3 [B,32,64,32]; UP: z [B,16,32,64]
  const __nv_bfloat16* skip;    // UP: x4 [B,16,32,64]
  const unsigned char* wstream; // the stage weight images back to back (ALL: the DOWN stages')
  const float* bias;
  __nv_bfloat16* y;             // DOWN: x4 [B,16,32,64]; UP / ALL: [B,32,64,32]
  // MODE_ALL only
  __nv_bfloat16* mid;           // x4 [B,16,32,64]: written by the DOWN stages, read by level 5 and (skip) by the UP stages
  __nv_bfloat16* z;             // [B,16,32,64]: level 5's output, the UP stages' input
  const unsigned char* wstream5; const float* bias5;        // level 5, per-rank form (fn_level5_desc.split)
  const unsigned char* wstream_up; const float* bias_up;    // the UP stages
  int* status;
  long long* timing;            // debug: 16 clock64 stamps per CTA (fn_debug_set_timing; nullptr in production)
  int timing_ctas;
  int B;
};

__device__ __forceinline__ void worker_barrier() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
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
__device__ __forceinline__ uint4 pack8(const float (&o)[8]) {
  bf16x8 st;
#pragma unroll
  for (int i = 0; i < 8; ++i) st.h[i] = __float2bfloat16(o[i]);
  return st.v;
}
// ---- cluster plumbing (as conv_l5c.cu)
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_smem_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void st_cluster128(uint32_t cluster_addr, const uint4& v) {
  asm volatile("st.shared::cluster.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(cluster_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok != 0;
}
static __device__ __noinline__ bool mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return true;
  const long long t0 = clock64();
  for (;;) {
#pragma unroll 1
    for (int i = 0; i < 32; ++i)
      if (mbar_try_wait_cluster(bar, parity)) return true;
    if (clock64() - t0 > (1ll << 31)) return false;
  }
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1) k_level4(const __grid_constant__ Params p) {
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

  if (warp == NWORK) {
    ptx::tmem_alloc(sbase + 256, 256u);
    ptx::tmem_relinquish();
    if (lane == 0) {
      for (int s = 0; s < NST; ++s) {
        ptx::mbar_init(bar(SB_FULL + s), 1);
        ptx::mbar_init(bar(SB_EMPTY + s), 2);          // one commit per issuing warp
      }
      ptx::mbar_init(bar(SB_TF), 2);
      ptx::mbar_init(bar(SB_XR), 1);
      ptx::mbar_init(bar(SB_XW), NWORK);
      if (MODE == MODE_ALL) l5c::init_barriers(bar(BAR5), NWORK);
      ptx::fence_mbar_init();
    }
    __syncwarp();
  }
  // biases of one half into shared memory, the v halves of the gated convs pre-halved: sigmoid(z) = 0.5 + 0.5 tanh(z / 2)
  auto load_biases = [&](const float* bias, bool down) {
    const int n = down ? NBIAS_DOWN : NBIAS_UP;
    for (int i = tid; i < n; i += THREADS) {
      float bv = __ldg(bias + i);
      const bool vhalf = down ? ((i >= B2 + 64 && i < B3) || i >= B4 + 64) : (i >= B6 + 64 && i < B7);
      if (vhalf) bv *= 0.5f;
      sbias[i] = bv;
    }
  };
  load_biases(p.bias, MODE != MODE_UP);
  // MODE_ALL: between two halves every thread of both CTAs passes here -- global writes of the finished half visible to the
  // cluster, the next half's biases in place, nobody still reading the old ones
  auto next_half = [&](auto load_next) {
    __threadfence();
    ptx::tc_fence_before_sync();
    __syncthreads();
    load_next();
    cluster_sync_all();
    ptx::tc_fence_after_sync();
  };
  ptx::tc_fence_before_sync();
  cluster_sync_all();                 // both CTAs' barriers exist before anything arrives on them
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  bool ok = true;
  long long* stamps = (MODE != MODE_ALL && p.timing != nullptr && (int)blockIdx.x < p.timing_ctas && tid == 0) ? p.timing + (size_t)blockIdx.x * 16 : nullptr;
  int nstamp = 0;
  auto stamp = [&]() { if (stamps) stamps[nstamp++] = clock64(); };
  stamp();
  // MODE_ALL: the level-5 stages' view of this launch (their own barrier set, the same TMEM columns and ring region)
  l5c::Ctx c5;
  c5.sbase = sbase; c5.bar0 = bar(BAR5); c5.sb = sb; c5.tmem_base = tmem_base;
  c5.warp = warp; c5.lane = lane; c5.tid = tid; c5.b = b; c5.rank = rank;
  c5.x = p.mid; c5.y = p.z; c5.yp = nullptr; c5.wstream = p.wstream5; c5.status = p.status;
  c5.stamps = nullptr; c5.nstamp = 0;
  auto load_biases5 = [&]() { l5c::load_biases(sbias, p.bias5, rank, tid, THREADS); };
  auto load_biases_up = [&]() { load_biases(p.bias_up, false); };

  if (warp == NWORK) {
    // =========================== producer: the weight stream (both CTAs read all of it) ===========================
    if (lane == 0) {
      const uint64_t pol = ptx::policy_evict_last();
      int gc = 0;                       // ring chunk counter: runs on from the DOWN stages to the UP stages in MODE_ALL
      auto stream = [&](const unsigned char* ws, bool down) {
        const int nstage = down ? 4 : 3;
        const int stage_bytes[4] = {down ? W1 : W5, down ? W2 : W6, down ? W3 : W7, W4};
        size_t off = 0;
        for (int st = 0; st < nstage && ok; ++st) {
          for (int rem = stage_bytes[st]; rem > 0 && ok; ++gc) {
            const int bytes = rem < CHUNK ? rem : CHUNK;
            const int s = gc % NST;
            if (gc >= NST) {
              ok = ptx::mbar_wait(bar(SB_EMPTY + s), (uint32_t)(((gc / NST) - 1) & 1));
              if (!ok) { atomicCAS(p.status, 0, 81); break; }
            }
            ptx::mbar_arrive_expect_tx(bar(SB_FULL + s), (uint32_t)bytes);
            ptx::bulk_g2s_hint(sbase + OFF_RING + s * CHUNK, ws + off, (uint32_t)bytes, bar(SB_FULL + s), pol);
            off += (size_t)bytes;
            rem -= bytes;
          }
        }
      };
      stream(p.wstream, MODE != MODE_UP);
    }
    __syncwarp();
    if constexpr (MODE == MODE_ALL) {
      // the ring region changes hands only when the finished half has consumed all of its chunks: both cluster barriers of
      // a hand-over are passed by this warp too
      next_half(load_biases5);
      l5c::producer(c5, ok);
      next_half(load_biases_up);
      if (lane == 0) {
        const uint64_t pol = ptx::policy_evict_last();
        int gc = NCHUNK_DOWN;
        const int stage_bytes[3] = {W5, W6, W7};
        size_t off = 0;
        for (int st = 0; st < 3 && ok; ++st) {
          for (int rem = stage_bytes[st]; rem > 0 && ok; ++gc) {
            const int bytes = rem < CHUNK ? rem : CHUNK;
            const int s = gc % NST;
            ok = ptx::mbar_wait(bar(SB_EMPTY + s), (uint32_t)(((gc / NST) - 1) & 1));
            if (!ok) { atomicCAS(p.status, 0, 86); break; }
            ptx::mbar_arrive_expect_tx(bar(SB_FULL + s), (uint32_t)bytes);
            ptx::bulk_g2s_hint(sbase + OFF_RING + s * CHUNK, p.wstream_up + off, (uint32_t)bytes, bar(SB_FULL + s), pol);
            off += (size_t)bytes;
            rem -= bytes;
          }
        }
      }
      __syncwarp();
    }
  } else {
    // =========================== workers ===========================
    const uint32_t leader = ptx::elect_one();
    const int q = warp & 3, jt = (warp >> 2) & 1, chalf = warp >> 3;
    const int r = q * 32 + lane;                      // accumulator row == TMEM lane
    const int row = r >> 3, col = jt * 8 + (r & 7);   // this thread's pixel inside the CTA's 16 x 16 part
    const int gcol = (int)rank * CW + col;            // column inside the image
    const int pix = jt * 128 + r;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t A = sbase + OFF_A;
    const uint32_t A_peer = map_to_cta(A, peer);
    const uint32_t xr_peer = map_to_cta(bar(SB_XR), peer), xw_peer = map_to_cta(bar(SB_XW), peer);
    const uint32_t pp16 = (uint32_t)((row + 1) * PU + (col + 1)) * 16u;                 // this pixel inside a plane
    // the CTA's border column is the peer's halo: rank 0's column 15 -> the peer's plane column 0, rank 1's column 0 -> 17
    const bool border = rank == 0 ? col == CW - 1 : col == 0;
    const uint32_t pp16_peer = (uint32_t)((row + 1) * PU + (rank == 0 ? 0 : PU - 1)) * 16u;
    uint32_t tf_par = 0, xr_par = 0, xw_par = 0;

    auto chunk_ready = [&](int gc, uint32_t lbo) -> uint64_t {
      const int s = gc % NST;
      ok = ok && ptx::mbar_wait(bar(SB_FULL + s), (uint32_t)((gc / NST) & 1));
      if (!ok) atomicCAS(p.status, 0, 82);
      ptx::tc_fence_after_sync();
      return ptx::smem_desc(sbase + OFF_RING + s * CHUNK, lbo, 128);
    };
    // everybody: this stage's accumulators are complete; the peer learns that this CTA's planes may be overwritten
    // readers_first: other warps read the operand planes during this stage's MMAs (the shortcut recovered from them), so
    // nobody -- neither this CTA's epilogue nor the peer -- may overwrite them before every warp has passed this point
    auto stage_done = [&](bool readers_first = false) {
      ok = ok && ptx::mbar_wait(bar(SB_TF), tf_par);
      if (!ok) atomicCAS(p.status, 0, 83);
      tf_par ^= 1u;
      ptx::tc_fence_after_sync();
      if (readers_first) worker_barrier();
      if (tid == 0) mbar_arrive_remote(xr_peer);
    };
    // before the first remote write of an epilogue: the peer's MMAs no longer read its planes
    auto peer_planes_free = [&]() {
      ok = ok && mbar_wait_cluster(bar(SB_XR), xr_par);
      if (!ok) atomicCAS(p.status, 0, 84);
      xr_par ^= 1u;
    };
    // after an epilogue: own writes (local + remote) visible to the tensor-core proxies, the peer told; then wait for the
    // local workers and for the peer's sixteen warps
    auto planes_exchanged = [&]() {
      fence_proxy_async_all();
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(xw_peer);
      worker_barrier();
      ok = ok && mbar_wait_cluster(bar(SB_XW), xw_par);
      if (!ok) atomicCAS(p.status, 0, 85);
      fence_proxy_async_all();            // the peer's generic-proxy writes, acquired above, before this CTA's tensor-core reads
      xw_par ^= 1u;
    };
    auto put_plane = [&](int plane, const uint4& v) {
      sts128(A + (uint32_t)(plane * PPAD * 16) + pp16, v);
      if (border) st_cluster128(A_peer + (uint32_t)(plane * PPAD * 16) + pp16_peer, v);
    };

    // ---- MMAs of one convolution whose operand sits in planes: NTAPS taps of KT k-steps.  Warps 0 and 1 each issue the
    // program of ONE of the CTA's two tiles (jw = warp): a single thread needs ~53 clocks per tcgen05.mma plus ~270 per
    // ring chunk (wait, commit, loop), more than the 64 clocks an N = 128 MMA executes -- with one issuer the stages ran at
    // 82-87 clocks per MMA (tools/phase_timing_l4.py); with two the ring chunk is consumed at the tensor pipe's pace.
    // tap_off(tap): byte offset of the tap's first k-step inside the A operand; kstride: bytes between k-steps;
    // first: the first MMA overwrites the accumulators; g0: global chunk index of the stage's first chunk
    auto issue_conv = [&](auto nconst, auto ktconst, auto ntconst, uint64_t a_tmpl0, uint32_t kstride, auto tap_off, int g0, bool first) {
      const uint64_t a_tmpl = a_tmpl0 + (uint64_t)((uint32_t)warp * 128u >> 4);      // this warp's tile
      const uint32_t d_tmem = tmem_base + (uint32_t)(warp * decltype(nconst)::value);
      constexpr int N = decltype(nconst)::value, KT = decltype(ktconst)::value, NTAPS = decltype(ntconst)::value;
      constexpr int SPC = CHUNK / (N * 32);            // k-steps per chunk
      constexpr uint32_t idesc = ptx::idesc_bf16_f32(128, N);
      if constexpr (SPC >= KT) {
        constexpr int TPC = SPC / KT;                  // taps per chunk
        constexpr int NCH = (NTAPS + TPC - 1) / TPC;
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
          const uint64_t b_chunk = chunk_ready(g0 + c, N * 16);
#pragma unroll
          for (int t = 0; t < TPC; ++t) {
            const int tap = c * TPC + t;
            if (tap < NTAPS) {
              const uint64_t a_tap = a_tmpl + (uint64_t)(tap_off(tap) >> 4);
#pragma unroll
              for (int kk = 0; kk < KT; ++kk) {
                if (leader) ptx::mma_bf16_ss(d_tmem, a_tap + (uint64_t)((kk * kstride) >> 4), b_chunk + (uint32_t)(((t * KT + kk) * N * 32) >> 4), idesc,
                                             (first && tap == 0 && kk == 0) ? 0u : 1u);
              }
            }
          }
          if (leader) ptx::mma_commit(bar(SB_EMPTY + (g0 + c) % NST));
        }
      } else {
        constexpr int CPT = KT / SPC;                  // chunks per tap
#pragma unroll 1
        for (int c = 0; c < NTAPS * CPT; ++c) {
          long long* ist = (stamps && c >= 8 && c < 10) ? stamps + 8 + 4 * (c - 8) : nullptr;     // debug: issuer phases of two chunks
          if (ist) ist[0] = clock64();
          const uint64_t b_chunk = chunk_ready(g0 + c, N * 16);
          if (ist) ist[1] = clock64();
          const int tap = c / CPT, kk0 = (c - tap * CPT) * SPC;
          const uint64_t a_tap = a_tmpl + (uint64_t)((tap_off(tap) + (uint32_t)kk0 * kstride) >> 4);
#pragma unroll
          for (int i = 0; i < SPC; ++i) {
            if (leader) ptx::mma_bf16_ss(d_tmem, a_tap + (uint64_t)((i * kstride) >> 4), b_chunk + (uint32_t)((i * N * 32) >> 4), idesc,
                                         (first && c == 0 && i == 0) ? 0u : 1u);
          }
          if (ist) ist[2] = clock64();
          if (leader) ptx::mma_commit(bar(SB_EMPTY + (g0 + c) % NST));
          if (ist) ist[3] = clock64();
        }
      }
    };
    auto stage_commit = [&]() {
      if (leader) ptx::mma_commit(bar(SB_TF));
      __syncwarp();
    };
    using I4 = std::integral_constant<int, 4>;
    using I8 = std::integral_constant<int, 8>;
    using I9 = std::integral_constant<int, 9>;
    using I1 = std::integral_constant<int, 1>;
    using N64 = std::integral_constant<int, 64>;
    using N128 = std::integral_constant<int, 128>;
    const uint64_t a_planes = ptx::smem_desc(A, PPAD * 16, PU * 16);
    auto tap_s1 = [](int tap) -> uint32_t {            // stride-1 3x3 on the planes
      const int di = tap / 3, dj = tap - 3 * di;
      return (uint32_t)((di * PU + dj) * 16);
    };

    // ---- epilogue of a plain conv (64 channels): acc + bias -> bf16 -> [elu | elu(-.)] planes
    auto epilogue_plain = [&](int boff) {
#pragma unroll 1
      for (int g = 0; g < 4; g += 2) {
        const int c = (chalf * 4 + g) * 8;
        float o[2][8], bv[2][8];
        lds_f8(sb + 4u * (uint32_t)(boff + c), bv[0]);
        lds_f8(sb + 4u * (uint32_t)(boff + c + 8), bv[1]);
        ptx::tmem_ld8x2(lane_base + (uint32_t)(jt * 64 + c), lane_base + (uint32_t)(jt * 64 + c + 8), o[0], o[1]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int i = 0; i < 8; ++i) o[h][i] += bv[h][i];
          uint4 o0, o1;
          concat_elu8(pack8(o[h]), o0, o1);
          put_plane((c >> 3) + h, o0);
          put_plane(8 + (c >> 3) + h, o1);
        }
      }
    };
    // gate of 8 channels of tile jt: (a + bu) * sigmoid(v + bv), product not yet rounded
    auto gate8 = [&](int boff, int c, float (&a)[8], float (&sg)[8]) {
      float gt[8], bu[8], bv[8];
      lds_f8(sb + 4u * (uint32_t)(boff + c), bu);
      lds_f8(sb + 4u * (uint32_t)(boff + 64 + c), bv);
      ptx::tmem_ld8x2(lane_base + (uint32_t)(jt * 128 + c), lane_base + (uint32_t)(jt * 128 + 64 + c), a, gt);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a[i] += bu[i];
        sg[i] = fmaf(tanh_approx(fmaf(gt[i], 0.5f, bv[i])), 0.5f, 0.5f);
      }
    };

    constexpr int GB = MODE == MODE_ALL ? NCHUNK_DOWN : 0;      // ring chunk index of the UP stages' first chunk
    if constexpr (MODE != MODE_UP) {
      // ---- stage 1 operand: x3 -> concat_elu -> parity sub-planes (row 32 / column 64 are TF's zero padding); each CTA
      // reads the 33 columns its 16 output columns touch
      {
        const __nv_bfloat16* xb = p.x + (size_t)b * LH3 * LW3 * C3;
        const int c0 = (int)rank * 2 * CW;
        constexpr int ITEMS = 33 * 33 * 4;
        constexpr int UN = (ITEMS + WTHREADS - 1) / WTHREADS;       // every load of a thread in flight at once
        for (int base = tid; base < ITEMS; base += WTHREADS * UN) {
          uint4 in[UN];
#pragma unroll
          for (int k = 0; k < UN; ++k) {
            const int idx = base + WTHREADS * k;
            const int ch = idx & 3, sp = idx >> 2;
            const int h = sp / 33, w = sp - h * 33;
            in[k] = make_uint4(0u, 0u, 0u, 0u);
            if (idx < ITEMS && h < LH3 && c0 + w < LW3) in[k] = __ldg(reinterpret_cast<const uint4*>(xb + ((size_t)h * LW3 + c0 + w) * C3 + ch * 8));
          }
#pragma unroll
          for (int k = 0; k < UN; ++k) {
            const int idx = base + WTHREADS * k;
            if (idx >= ITEMS) break;
            const int ch = idx & 3, sp = idx >> 2;
            const int h = sp / 33, w = sp - h * 33;
            uint4 o0, o1;
            concat_elu8(in[k], o0, o1);
            const uint32_t d0 = A + (uint32_t)(((h & 1) * 2 + (w & 1)) * SUB1_BYTES) + (uint32_t)(ch * P1 + (h >> 1) * PS1 + (w >> 1)) * 16u;
            sts128(d0, o0);
            sts128(d0 + (uint32_t)(4 * P1 * 16), o1);
          }
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before_sync();
      worker_barrier();
      stamp();

      // ---- stage 1: stride-2 conv_1 of resnet_4_downsample, K = 9 x 64, N = 64
      if (warp < 2) {
        ptx::tc_fence_after_sync();
        auto tap_s2 = [](int tap) -> uint32_t {
          const int di = tap / 3, dj = tap - 3 * di;
          return (uint32_t)(((di & 1) * 2 + (dj & 1)) * SUB1_BYTES + ((di >> 1) * PS1 + (dj >> 1)) * 16);
        };
        issue_conv(N64{}, I4{}, I9{}, ptx::smem_desc(A, P1 * 16, PS1 * 16), 2u * P1 * 16u, tap_s2, G1, true);
        stage_commit();
      }
      // the shortcut of stage 2, front channel pad (reference :158-163): output channels 32.. (the warps with chalf = 1, none of
      // which issues) take the 2x2 average pool of x3.  The four input pixels of output pixel (row, col) are entry
      // [row][col] of the four parity sub-planes, and x is exactly recoverable from [elu(x) | elu(-x)] (x > 0: elu(x) = x;
      // x <= 0: elu(-x) = -x), so the pool is computed from shared memory while stage 1's MMAs read the same planes --
      // no global loads (scattered 16-byte loads during a stage's MMAs delayed the weight ring by 3.5 k clocks)
      float pool[4][8];
      if (chalf == 1) {
        const uint32_t e0 = A + (uint32_t)(row * PS1 + col) * 16u;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
#pragma unroll
          for (int i = 0; i < 8; ++i) pool[g][i] = 0.f;
#pragma unroll
          for (int sp = 0; sp < 4; ++sp) {
            bf16x8 pos, neg;
            pos.v = lds128(e0 + (uint32_t)(sp * SUB1_BYTES + g * P1 * 16));
            neg.v = lds128(e0 + (uint32_t)(sp * SUB1_BYTES + (4 + g) * P1 * 16));
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float fp = __bfloat162float(pos.h[i]);
              pool[g][i] += fp > 0.f ? fp : -__bfloat162float(neg.h[i]);
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) pool[g][i] *= 0.25f;
        }
      }
      stage_done(true);
      stamp();
      // the zero border of the planes (the region held the sub-planes until now): rows 0 / 17 and the image-edge column;
      // the halo column towards the peer is written by the peer
      for (int idx = tid; idx < 16 * (2 * PU + PV - 2); idx += WTHREADS) {
        const int pl = idx / (2 * PU + PV - 2), e = idx - pl * (2 * PU + PV - 2);
        int pr, pc;
        if (e < PU) { pr = 0; pc = e; }
        else if (e < 2 * PU) { pr = PV - 1; pc = e - PU; }
        else { pr = e - 2 * PU + 1; pc = rank == 0 ? 0 : PU - 1; }
        sts128(A + (uint32_t)(pl * PPAD + pr * PU + pc) * 16u, make_uint4(0u, 0u, 0u, 0u));
      }
      peer_planes_free();
      epilogue_plain(B1);
      planes_exchanged();
      stamp();

      // ---- stage 2: conv_2 of the down block, gate, shortcut = 2x2 average pool of x3 in the LAST 32 channels
      if (warp < 2) {
        ptx::tc_fence_after_sync();
        issue_conv(N128{}, I8{}, I9{}, a_planes, 2u * PPAD * 16u, tap_s1, G2, true);
        stage_commit();
      }
      stage_done();
      stamp();
      peer_planes_free();
      {
        const uint32_t r1_row = sbase + OFF_R1 + (uint32_t)pix * 128u;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int c = (chalf * 4 + g) * 8;
          float a[8], sg[8], o[8];
          gate8(B2, c, a, sg);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = a[i] * sg[i];
          if (chalf == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] += pool[g][i];
          }
          const uint4 raw = pack8(o);
          sts128(r1_row + (uint32_t)(((c >> 3) ^ (pix & 7)) * 16), raw);             // kept for stage 4's shortcut (same thread)
          uint4 o0, o1;
          concat_elu8(raw, o0, o1);
          put_plane(c >> 3, o0);
          put_plane(8 + (c >> 3), o1);
        }
      }
      planes_exchanged();
      stamp();

      // ---- stage 3: conv_1 of resnet_4_0
      if (warp < 2) {
        ptx::tc_fence_after_sync();
        issue_conv(N64{}, I8{}, I9{}, a_planes, 2u * PPAD * 16u, tap_s1, G3, true);
        stage_commit();
      }
      stage_done();
      stamp();
      peer_planes_free();
      epilogue_plain(B3);
      planes_exchanged();
      stamp();

      // ---- stage 4: conv_2 of resnet_4_0, gate, shortcut = the block input kept in shared memory -> x4 in global memory
      if (warp < 2) {
        ptx::tc_fence_after_sync();
        issue_conv(N128{}, I8{}, I9{}, a_planes, 2u * PPAD * 16u, tap_s1, G4, true);
        stage_commit();
      }
      stage_done();
      stamp();
      {
        const uint32_t r1_row = sbase + OFF_R1 + (uint32_t)pix * 128u;
        __nv_bfloat16* dst = (MODE == MODE_ALL ? p.mid : p.y) + (((size_t)b * IH + row) * IW + gcol) * C4;
#pragma unroll 1
        for (int g = 0; g < 4; ++g) {
          const int c = (chalf * 4 + g) * 8;
          float a[8], sg[8], o[8];
          gate8(B4, c, a, sg);
          bf16x8 rv;
          rv.v = lds128(r1_row + (uint32_t)(((c >> 3) ^ (pix & 7)) * 16));
          // product and shortcut add rounded separately, as in the layer-by-layer kernels
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = __fadd_rn(__fmul_rn(a[i], sg[i]), __bfloat162float(rv.h[i]));
          if (ok) *reinterpret_cast<uint4*>(dst + c) = pack8(o);
        }
      }
    }
    if constexpr (MODE == MODE_ALL) {
      peer_planes_free();               // consumes the peer's arrival for stage 4, so that XR's phase matches the UP stages' count
      next_half(load_biases5);
      l5c::workers<2, true, NWORK>(c5, ok);
      next_half(load_biases_up);
    }
    if constexpr (MODE != MODE_DOWN) {
      // ---- stage 5 operands from global memory: planes of concat_elu(z) with the halo column read directly (zero outside
      // the image), centre-only planes of concat_elu(a) for the nin
      {
        const __nv_bfloat16* zb = (MODE == MODE_ALL ? p.z : p.x) + (size_t)b * IH * IW * C4;
        const __nv_bfloat16* ab = (MODE == MODE_ALL ? p.mid : p.skip) + (size_t)b * IH * IW * C4;
        constexpr int ITEMS_Z = PV * PU * 8, ITEMS = ITEMS_Z + PS * 8;
        constexpr int UN = (ITEMS + WTHREADS - 1) / WTHREADS;       // every load of a thread in flight at once
        for (int base = tid; base < ITEMS; base += WTHREADS * UN) {
          uint4 in[UN];
#pragma unroll
          for (int k = 0; k < UN; ++k) {
            const int idx = base + WTHREADS * k;
            in[k] = make_uint4(0u, 0u, 0u, 0u);
            if (idx < ITEMS_Z) {
              const int ch = idx & 7, sp = idx >> 3;
              const int pr = sp / PU, pc = sp - pr * PU;
              const int h = pr - 1, w = (int)rank * CW + pc - 1;
              if (h >= 0 && h < IH && w >= 0 && w < IW) in[k] = l5c::ld_in<MODE == MODE_ALL>(reinterpret_cast<const uint4*>(zb + ((size_t)h * IW + w) * C4 + ch * 8));
            } else if (idx < ITEMS) {
              const int i2 = idx - ITEMS_Z;
              const int ch = i2 & 7, sp = i2 >> 3;
              const int h = sp >> 4, w = (int)rank * CW + (sp & 15);
              in[k] = l5c::ld_in<MODE == MODE_ALL>(reinterpret_cast<const uint4*>(ab + ((size_t)h * IW + w) * C4 + ch * 8));
            }
          }
#pragma unroll
          for (int k = 0; k < UN; ++k) {
            const int idx = base + WTHREADS * k;
            if (idx >= ITEMS) break;
            uint4 o0, o1;
            concat_elu8(in[k], o0, o1);
            if (idx < ITEMS_Z) {
              const int ch = idx & 7, sp = idx >> 3;
              const uint32_t d0 = A + (uint32_t)(ch * PPAD + sp) * 16u;
              sts128(d0, o0);
              sts128(d0 + (uint32_t)(8 * PPAD * 16), o1);
            } else {
              const int i2 = idx - ITEMS_Z;
              const int ch = i2 & 7, sp = i2 >> 3;
              const uint32_t d0 = sbase + OFF_S + (uint32_t)(ch * PSPAD + sp) * 16u;
              sts128(d0, o0);
              sts128(d0 + (uint32_t)(8 * PSPAD * 16), o1);
            }
          }
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before_sync();
      worker_barrier();
      stamp();

      // ---- stage 5: conv_1 of resnet_up_1_0 on concat_elu(z) plus the nin on concat_elu(a) (8 more k-steps, centre tap)
      if (warp < 2) {
        ptx::tc_fence_after_sync();
        issue_conv(N64{}, I8{}, I9{}, a_planes, 2u * PPAD * 16u, tap_s1, GB + G5, true);
        auto tap_c = [](int) -> uint32_t { return 0u; };
        issue_conv(N64{}, I8{}, I1{}, ptx::smem_desc(sbase + OFF_S, PSPAD * 16, CW * 16), 2u * PSPAD * 16u, tap_c, GB + G5 + 9, false);
        stage_commit();
      }
      // the shortcut of stage 6 (z, this thread's channels), recovered from the [elu(z) | elu(-z)] planes while stage 5 reads them
      bf16x8 zres[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        bf16x8 pos, neg;
        pos.v = lds128(A + (uint32_t)((chalf * 4 + g) * PPAD * 16) + pp16);
        neg.v = lds128(A + (uint32_t)((8 + chalf * 4 + g) * PPAD * 16) + pp16);
#pragma unroll
        for (int i = 0; i < 8; ++i) zres[g].h[i] = __bfloat162float(pos.h[i]) > 0.f ? pos.h[i] : __hneg(neg.h[i]);
      }
      stage_done();
      stamp();
      peer_planes_free();
      epilogue_plain(B5);
      planes_exchanged();
      stamp();

      // ---- stage 6: conv_2, gate, shortcut = z (global, L2-resident); the result feeds the transposed conv RAW (no
      // nonlinearity, reference :219): 8 planes
      if (warp < 2) {
        ptx::tc_fence_after_sync();
        issue_conv(N128{}, I8{}, I9{}, a_planes, 2u * PPAD * 16u, tap_s1, GB + G6, true);
        stage_commit();
      }
      stage_done();
      stamp();
      peer_planes_free();
      {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int c = (chalf * 4 + g) * 8;
          float a[8], sg[8], o[8];
          gate8(B6, c, a, sg);
          const bf16x8& rv = zres[g];
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = __fadd_rn(__fmul_rn(a[i], sg[i]), __bfloat162float(rv.h[i]));
          put_plane(c >> 3, pack8(o));
        }
      }
      planes_exchanged();
      stamp();

      // ---- stage 7: transposed conv as four output phases, K = 64 per tap, N = 32 per phase; a 16 KB
      // chunk = 16 k-steps = four taps.  Accumulators: tile j, phase ph at columns j * 128 + ph * 32
      if (warp < 2) {
        ptx::tc_fence_after_sync();
        constexpr uint32_t idesc = ptx::idesc_bf16_f32(128, 32);
#pragma unroll 1
        for (int c = 0; c < chunks_of(W7); ++c) {
          const uint64_t b_chunk = chunk_ready(GB + G7 + c, 32 * 16);
#pragma unroll
          for (int h4 = 0; h4 < 4; ++h4) {
            const int tap = 4 * c + h4;
            if (tap < 9) {
              const int di = tap / 3, dj = tap - 3 * di;
              // out[2m + p] takes tap d = p from in[m] and, for p = 0, tap d = 2 from in[m - 1]
              const uint64_t a_tap = a_planes + (uint64_t)((uint32_t)((((di == 2) ? 0 : 1) * PU + ((dj == 2) ? 0 : 1)) * 16) >> 4);
              const uint32_t d_ph = (uint32_t)((((di == 1) ? 2 : 0) + ((dj == 1) ? 1 : 0)) * 32);
              const bool first_tap = tap == 0 || tap == 1 || tap == 3 || tap == 4;
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                if (leader) ptx::mma_bf16_ss(tmem_base + (uint32_t)(warp * 128) + d_ph, a_tap + (uint64_t)((kk * 2u * PPAD * 16u + (uint32_t)warp * 128u) >> 4),
                                             b_chunk + (uint32_t)(((h4 * 4 + kk) * 32 * 32) >> 4), idesc, (first_tap && kk == 0) ? 0u : 1u);
              }
            }
          }
          if (leader) ptx::mma_commit(bar(SB_EMPTY + (GB + G7 + c) % NST));
        }
        stage_commit();
      }
      stage_done();
      stamp();
      {
        __nv_bfloat16* yb = p.y + (size_t)b * LH3 * LW3 * C3;
#pragma unroll
        for (int ph = 0; ph < 4; ++ph) {
          const int oy = 2 * row + (ph >> 1), ox = 2 * gcol + (ph & 1);
          __nv_bfloat16* dst = yb + ((size_t)oy * LW3 + ox) * C3;
          const int c = chalf * 16;
          float o[2][8], bv[2][8];
          lds_f8(sb + 4u * (uint32_t)(B7 + c), bv[0]);
          lds_f8(sb + 4u * (uint32_t)(B7 + c + 8), bv[1]);
          ptx::tmem_ld8x2(lane_base + (uint32_t)(jt * 128 + ph * 32 + c), lane_base + (uint32_t)(jt * 128 + ph * 32 + c + 8), o[0], o[1]);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int i = 0; i < 8; ++i) o[h][i] += bv[h][i];
            if (ok) *reinterpret_cast<uint4*>(dst + c + 8 * h) = pack8(o[h]);
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
    ptx::tmem_dealloc(tmem_base, 256u);
  }
}

}  // namespace l4

int* tc_status_word();   // conv_tc.cu
void tc_timing_buffer(long long** buf, int* ctas);   // conv_tc.cu: debug stamp buffer (fn_debug_set_timing)

int level4_fwd(const fn_level4_desc& d, cudaStream_t stream) {
  using namespace l4;
  Params p;
  memset(&p, 0, sizeof(p));
  p.x = (const __nv_bfloat16*)d.x;
  p.skip = (const __nv_bfloat16*)d.skip;
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
    cudaError_t e = cudaFuncSetAttribute(k_level4<MODE_DOWN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DOWN);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_level4<MODE_UP>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_UP);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(level4): %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  if (d.up) k_level4<MODE_UP><<<2 * d.B, THREADS, SMEM_UP, stream>>>(p);
  else k_level4<MODE_DOWN><<<2 * d.B, THREADS, SMEM_DOWN, stream>>>(p);
  count_launch();
  note_tcgen05(true);
  return check_launch("k_level4");
}

int level45_fwd(const fn_level45_desc& d, cudaStream_t stream) {
  using namespace l4;
  Params p;
  memset(&p, 0, sizeof(p));
  p.x = (const __nv_bfloat16*)d.x;
  p.wstream = (const unsigned char*)d.w_down;
  p.bias = d.bias_down;
  p.y = (__nv_bfloat16*)d.y;
  p.mid = (__nv_bfloat16*)d.x4;
  p.z = (__nv_bfloat16*)d.z;
  p.wstream5 = (const unsigned char*)d.w_level5;
  p.bias5 = d.bias_level5;
  p.wstream_up = (const unsigned char*)d.w_up;
  p.bias_up = d.bias_up;
  p.status = tc_status_word();
  if (p.status == nullptr) return (int)cudaErrorMemoryAllocation;
  p.B = d.B;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_level4<MODE_ALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ALL);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(level45): %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  k_level4<MODE_ALL><<<2 * d.B, THREADS, SMEM_ALL, stream>>>(p);
  count_launch();
  note_tcgen05(true);
  return check_launch("k_level45");
}

long long level4_weight_bytes(int up) {
  using namespace l4;
  auto r = [](int b) { return (long long)b; };
  return up ? r(W5) + r(W6) + r(W7) : r(W1) + r(W2) + r(W3) + r(W4);
}

}  // namespace fn
