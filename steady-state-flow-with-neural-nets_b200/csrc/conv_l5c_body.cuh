// The 8x16 level of flow_net in one launch, TWO CTAs per image (thread-block cluster of 2, output channels split).
//
// conv_l5.cu carries one image per CTA through resnet_5_downsample, resnet_5_0 and up_conv_1 (reference
// flow_architecture.py:207-215).  At the bench batch (64) that occupies 64 of the 148 SMs for ~60 k clocks of tcgen05.mma
// each (M = 128 is all an image has) while 3.4 MB of weights stream through every one of them: 62 us warm, 80 us inside
// the network -- the longest launch of the forward.  Here the two CTAs of a cluster share an image: each computes HALF of
// the output channels of every convolution (N = 64 / 128 / 64 / 128 / 4 x 32), streams only its half of the weights
// (1.7 MB) and, in the epilogue, writes its [elu | elu(-.)] planes of the next operand into BOTH CTAs' shared memory
// (st.shared::cluster), so that each CTA again holds the whole K = 256-channel operand.  Per stage:
//
//   own MMAs done  -> tell the peer (remote mbarrier arrive on its XR): "my planes may be overwritten"
//   epilogue       -> wait for the peer's XR, write planes locally and remotely, fence.proxy.async, per-warp remote arrive
//                     on the peer's XW
//   next MMAs      -> wait for the local worker barrier and the own XW (the peer's 8 warps have written)
//
// Same arithmetic, rounding points and MMA order per accumulator column as conv_l5.cu (bit-identical output).
#pragma once
#include <cuda.h>

#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"
#include "tma.cuh"

namespace fn {
using namespace tma;
namespace l5c {

constexpr int NWORK = 16;                      // worker warps of the stand-alone kernel (the stages take 8 or 16)
constexpr int THREADS = NWORK * 32 + 32;
constexpr int NST = 2, CHUNK = 32768;     // ring: two 32-KB slots during stage 1, THREE from stage 2 on (see Ring below)
constexpr int HDR = 4096;                      // [0,64) mbarriers, [256] TMEM base, [512, 512 + 416*4) biases
constexpr int C4 = 64, C5 = 128, IH = 16, IW = 32;
constexpr int CL = C5 / 2;                     // channels of a stage-1..4 output this CTA computes
// stage 1 operand: the four (row, column)-parity sub-planes of concat_elu(x4), each [16 planes][17 x][9 y][8]
constexpr int PU1 = 9, PV1 = 17, P1 = PU1 * PV1;
constexpr int SUB1_BYTES = 16 * P1 * 16;
// stages 2-5 operand: [32 planes][18 x][10 y][8] with a zero border (TF 'SAME'), plane stride padded to 1 (mod 8)
constexpr int PU = 10, PV = 18, P = PU * PV, PPAD = 185;
constexpr int PLANES_BYTES = 32 * PPAD * 16;
constexpr int OFF_RING = HDR;
constexpr int OFF_A = OFF_RING + NST * CHUNK;
constexpr int OFF_R1 = OFF_A + PLANES_BYTES;   // this CTA's half of resnet_5_downsample's raw output [128 px][64 ch], 16-byte units swizzled
constexpr int SMEM = OFF_A + 4 * SUB1_BYTES;
static_assert(OFF_R1 + 128 * 128 <= SMEM, "shortcut copy must fit behind the planes");
// per-rank biases, floats: conv_1 down, conv_2 down (u | v, v pre-halved), conv_1, conv_2, up_conv
constexpr int B1 = 0, B2 = 64, B3 = 192, B4 = 256, B5 = 384, NBIAS = 416;
// per-rank weight stream, bytes per stage
constexpr int W1 = 72 * 64 * 32, W2 = 144 * 128 * 32, W3 = 144 * 64 * 32, W4 = 144 * 128 * 32, W5 = 72 * 32 * 32;
constexpr long long WTOTAL = (long long)W1 + W2 + W3 + W4 + W5;
__host__ __device__ constexpr int chunks_of(int bytes) { return (bytes + CHUNK - 1) / CHUNK; }
constexpr int G1 = 0, G2 = G1 + chunks_of(W1), G3 = G2 + chunks_of(W2), G4 = G3 + chunks_of(W3), G5 = G4 + chunks_of(W4);

enum { SB_FULL = 0, SB_EMPTY = 2, SB_TF = 4, SB_XR = 5, SB_XW = 6, SB_FULL2 = 7, SB_EMPTY2 = 8 };

// Ring.  An N = 128 stage consumes 4 KB of weights per 64-clock MMA; one 32-KB slot takes ~1000 clocks to refill (round trip
// measured by tools/ring_bench.cu), so with two slots -- one in flight while the other is multiplied -- the stage ran at
// ~100 clocks per MMA.  Stage 1's parity sub-planes fill the shared memory; once its MMAs are done, 45 KB behind the
// stride-1 planes are free and become a third slot: chunk gc of the stream lives in slot ring_slot(gc), which it is the
// ring_fill(gc)-th chunk to occupy (slot 2's "zeroth occupant" are the sub-planes: its EMPTY barrier gets an arrival once
// stage 1's MMAs are complete and every warp has read its shortcut from them).
constexpr int OFF_SLOT2 = OFF_A + PLANES_BYTES + 128 * 128;
static_assert(OFF_SLOT2 + CHUNK <= SMEM && (OFF_SLOT2 % 128) == 0, "third ring slot");
__host__ __device__ constexpr int ring_slot(int gc) { return gc < G2 ? (gc & 1) : ((gc - G2) % 3 == 0 ? 1 : ((gc - G2) % 3 == 1 ? 2 : 0)); }
__host__ __device__ constexpr int ring_fill(int gc) {        // 0-based count of this slot's earlier chunks
  return gc < G2 ? (gc >> 1) : (gc - G2) / 3 + (ring_slot(gc) == 0 ? (G2 + 1) / 2 : ring_slot(gc) == 1 ? G2 / 2 : 0);
}
__host__ __device__ constexpr int ring_offset(int slot) { return slot < 2 ? OFF_RING + slot * CHUNK : OFF_SLOT2; }
__host__ __device__ constexpr int ring_full(int slot) { return slot < 2 ? SB_FULL + slot : SB_FULL2; }
__host__ __device__ constexpr int ring_empty(int slot) { return slot < 2 ? SB_EMPTY + slot : SB_EMPTY2; }

struct Params {
  const __nv_bfloat16* x;       // [B,16,32,64]
  const unsigned char* wstream; // rank 0's five weight images back to back, then rank 1's
  const float* bias;            // [2][416]
  __nv_bfloat16* y;             // [B,16,32,64]
  __nv_bfloat16* yp;            // optional planes form of y
  int* status;
  long long* timing;            // debug: 16 clock64 stamps per CTA (fn_debug_set_timing; nullptr in production)
  int timing_ctas;
  int B;
};

template <int ID, int NT>
__device__ __forceinline__ void worker_barrier() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(NT) : "memory"); }
// the input of the level: read-only for the whole launch (ld.global.nc), or written earlier in the SAME launch by this
// cluster (the merged levels-4+5 kernel): then through L2 (ld.global.cg), after a cluster barrier
template <bool COHERENT>
__device__ __forceinline__ uint4 ld_in(const uint4* p) { return COHERENT ? __ldcg(p) : __ldg(p); }
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
// ---- cluster plumbing
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


// What a phase of this kernel needs from the launch it runs in: conv_l5c.cu's own kernel, or the merged levels-4+5 kernel
// of conv_l4.cu, where the level-5 stages sit between the two halves of level 4.
struct Ctx {
  uint32_t sbase;               // shared window of the CTA
  uint32_t bar0;                // address of this phase's first mbarrier (SB_* indices are relative to it)
  uint32_t sb;                  // address of this rank's NBIAS biases (floats, v halves of the gated convs pre-halved)
  uint32_t tmem_base;           // 256 TMEM columns
  int warp, lane, tid, b;       // warps 0..7 are the workers; b = image
  uint32_t rank;
  const __nv_bfloat16* x;       // [B,16,32,64]
  __nv_bfloat16* y;             // [B,16,32,64]
  __nv_bfloat16* yp;            // optional planes form of y
  const unsigned char* wstream; // rank 0's stream, then rank 1's
  int* status;
  long long* stamps;            // debug clock stamps of thread 0 (nullptr in production)
  int nstamp;
};

__device__ __forceinline__ void init_barriers(uint32_t bar0, int nworkers) {        // one thread; nworkers: worker warps
  auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  for (int s = 0; s < 3; ++s) {
    ptx::mbar_init(bar(ring_full(s)), 1);
    ptx::mbar_init(bar(ring_empty(s)), 1);
  }
  ptx::mbar_init(bar(SB_TF), 2);        // warps 0 and 1 both commit a stage (see the K split below)
  ptx::mbar_init(bar(SB_XR), 1);
  ptx::mbar_init(bar(SB_XW), (uint32_t)nworkers);
}
// NBIAS floats of this rank into shared memory, the v halves of the gated convs pre-halved: sigmoid(z) = 0.5 + 0.5 tanh(z / 2)
__device__ __forceinline__ void load_biases(float* sbias, const float* bias, uint32_t rank, int tid, int nthreads) {
  for (int i = tid; i < NBIAS; i += nthreads) {
    float bv = __ldg(bias + rank * NBIAS + i);
    if ((i >= B2 + CL && i < B3) || (i >= B4 + CL && i < B5)) bv *= 0.5f;
    sbias[i] = bv;
  }
}

// the producer warp: this rank's weight stream through the ring
__device__ __forceinline__ void producer(Ctx& c, bool& ok) {
  const uint32_t sbase = c.sbase, rank = c.rank;
  const int lane = c.lane;
  auto bar = [&](int i) { return c.bar0 + 8u * (uint32_t)i; };
  struct { const unsigned char* wstream; int* status; } p = {c.wstream, c.status};
  // =========================== producer: this rank's weight stream ===========================
  if (lane == 0) {
    const unsigned char* ws = p.wstream + (size_t)rank * (size_t)WTOTAL;
    const int stage_bytes[5] = {W1, W2, W3, W4, W5};
    const uint64_t pol = ptx::policy_evict_last();
    int gc = 0;
    size_t off = 0;
    for (int st = 0; st < 5 && ok; ++st) {
      for (int rem = stage_bytes[st]; rem > 0 && ok; ++gc) {
        const int bytes = rem < CHUNK ? rem : CHUNK;
        const int s = ring_slot(gc), f = ring_fill(gc);
        if (f >= 1 || s == 2) {       // the slot's previous occupant has been multiplied (slot 2: stage 1's sub-planes are dead)
          ok = ptx::mbar_wait(bar(ring_empty(s)), (uint32_t)((s == 2 ? f : f - 1) & 1));
          if (!ok) { atomicCAS(p.status, 0, 71); break; }
        }
        ptx::mbar_arrive_expect_tx(bar(ring_full(s)), (uint32_t)bytes);
        ptx::bulk_g2s_hint(sbase + ring_offset(s), ws + off, (uint32_t)bytes, bar(ring_full(s)), pol);
        off += (size_t)bytes;
        rem -= bytes;
      }
    }
  }
  __syncwarp();
}

// warps 0..NW-1 (NW = 8 or 16): operand build, MMA issue (warp 0), epilogues -- warp w owns TMEM lane quadrant w & 3 and the
// (w >> 2)-th part of this CTA's output channels.  BARID: the named barrier of these NW * 32 threads.
template <int BARID, bool COHERENT, int NW>
__device__ __forceinline__ void workers(Ctx& c, bool& ok) {
  const uint32_t sbase = c.sbase, sb = c.sb, tmem_base = c.tmem_base, rank = c.rank, peer = c.rank ^ 1u;
  const int warp = c.warp, lane = c.lane, tid = c.tid, b = c.b;
  auto bar = [&](int i) { return c.bar0 + 8u * (uint32_t)i; };
  struct { const __nv_bfloat16* x; __nv_bfloat16* y; __nv_bfloat16* yp; int* status; } p = {c.x, c.y, c.yp, c.status};
  auto stamp = [&]() { if (c.stamps) c.stamps[c.nstamp++] = clock64(); };
  // =========================== workers ===========================
  const uint32_t leader = ptx::elect_one();
  constexpr int WT = NW * 32, NG = 32 / NW, NG5 = 16 / NW;      // worker threads; 8-channel groups per thread (stages 1-4 / stage 5)
  static_assert(NW == 8 || NW == 16, "8 or 16 worker warps");
  const int q = warp & 3, part = warp >> 2;
  const int r = q * 32 + lane;                      // accumulator row == TMEM lane
  const int py = r & 7, px = r >> 3;                // pixel (y, x) of the 8 x 16 image
  const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
  const uint32_t A = sbase + OFF_A;
  const uint32_t A_peer = map_to_cta(A, peer);
  const uint32_t xr_peer = map_to_cta(bar(SB_XR), peer), xw_peer = map_to_cta(bar(SB_XW), peer);
  const uint32_t pp16 = (uint32_t)((px + 1) * PU + (py + 1)) * 16u;        // this pixel inside a stage-2..5 plane
  const uint32_t r1_row = sbase + OFF_R1 + (uint32_t)r * 128u;
  uint32_t tf_par = 0, x_par = 0;

  auto chunk_ready = [&](int gc, uint32_t lbo) -> uint64_t {
    const int s = ring_slot(gc);
    ok = ok && ptx::mbar_wait(bar(ring_full(s)), (uint32_t)(ring_fill(gc) & 1));
    if (!ok) atomicCAS(p.status, 0, 72);
    ptx::tc_fence_after_sync();
    return ptx::smem_desc(sbase + ring_offset(s), lbo, 128);
  };
  // everybody: this stage's accumulators are complete; the peer learns that this CTA's planes may be overwritten
  // readers_first: other warps read the operand planes during this stage's MMAs (the shortcut recovered from them), so
  // nobody -- neither this CTA's epilogue nor the peer -- may overwrite them before every warp has passed this point
  auto stage_done = [&](bool readers_first = false) {
    ok = ok && ptx::mbar_wait(bar(SB_TF), tf_par);
    if (!ok) atomicCAS(p.status, 0, 73);
    tf_par ^= 1u;
    ptx::tc_fence_after_sync();
    if (readers_first) {
      worker_barrier<BARID, WT>();
      // stage 1 only: its sub-planes are dead now (MMAs complete, every warp has taken its shortcut from them) -- the memory
      // behind the stride-1 planes becomes the ring's third slot
      if (tid == 0) ptx::mbar_arrive(bar(SB_EMPTY2));
    }
    if (tid == 0) mbar_arrive_remote(xr_peer);
  };
  // before the first remote write of an epilogue: the peer's MMAs no longer read its planes
  auto peer_planes_free = [&]() {
    ok = ok && mbar_wait_cluster(bar(SB_XR), x_par);
    if (!ok) atomicCAS(p.status, 0, 74);
  };
  // after an epilogue: own writes (local + remote) visible to the tensor-core proxies, the peer told; then wait for the
  // local workers and for the peer's eight warps
  auto planes_exchanged = [&]() {
    fence_proxy_async_all();
    ptx::tc_fence_before_sync();
    __syncwarp();
    if (lane == 0) mbar_arrive_remote(xw_peer);
    worker_barrier<BARID, WT>();
    ok = ok && mbar_wait_cluster(bar(SB_XW), x_par);
    if (!ok) atomicCAS(p.status, 0, 75);
    fence_proxy_async_all();            // the peer's generic-proxy writes, acquired above, before this CTA's tensor-core reads
    x_par ^= 1u;
  };
  auto put_plane = [&](uint32_t off, const uint4& v) {       // the same 16 bytes into both CTAs' operand
    sts128(A + off, v);
    st_cluster128(A_peer + off, v);
  };

  // ---- stage 1 operand: x4 -> concat_elu -> parity sub-planes (row 16 / column 32 are TF's zero padding); each CTA
  // builds all of it (64 KB of global reads) -- cheaper than exchanging it
  {
    const __nv_bfloat16* xb = p.x + (size_t)b * IH * IW * C4;
    constexpr int ITEMS = (IH + 1) * (IW + 1) * 8;
    constexpr int UN = 9;                 // loads in flight per thread (two rounds)
    for (int base = tid; base < ITEMS; base += WT * UN) {
      uint4 in[UN];
#pragma unroll
      for (int k = 0; k < UN; ++k) {
        const int idx = base + WT * k;
        const int ch = idx & 7, sp = idx >> 3;
        const int h = sp / (IW + 1), w = sp - h * (IW + 1);
        in[k] = make_uint4(0u, 0u, 0u, 0u);
        if (idx < ITEMS && h < IH && w < IW) in[k] = ld_in<COHERENT>(reinterpret_cast<const uint4*>(xb + ((size_t)h * IW + w) * C4 + ch * 8));
      }
#pragma unroll
      for (int k = 0; k < UN; ++k) {
        const int idx = base + WT * k;
        if (idx >= ITEMS) break;
        const int ch = idx & 7, sp = idx >> 3;
        const int h = sp / (IW + 1), w = sp - h * (IW + 1);
        uint4 o0, o1;
        concat_elu8(in[k], o0, o1);
        const uint32_t d0 = A + (uint32_t)(((h & 1) * 2 + (w & 1)) * SUB1_BYTES) + (uint32_t)(ch * P1 + (w >> 1) * PU1 + (h >> 1)) * 16u;
        sts128(d0, o0);
        sts128(d0 + (uint32_t)(8 * P1 * 16), o1);
      }
    }
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  worker_barrier<BARID, WT>();
  stamp();

  // K split: ONE thread needs ~53 clocks per tcgen05.mma plus ~270 per ring chunk (wait, commit, loop) -- 82 clocks per N = 128 MMA
  // that executes in 64 (and 67 per N = 64 MMA that executes in 48).  An image is a single M = 128 tile here, so the split
  // is over K: warp 0 issues the even ring chunks of a stage into accumulator 0, warp 1 the odd ones into accumulator 1
  // (TMEM columns 128..), each chunk still consumed and released by exactly one of them; the epilogues add the two.
  const uint32_t acc_mine = tmem_base + (uint32_t)(warp & 1) * 128u;       // issuing warps only

  // ---- stage 1 MMAs: stride-2 conv, K = 9 x 128, this CTA's N = 64; a 32 KB chunk = 16 k-steps = two taps
  if (warp < 2) {
    ptx::tc_fence_after_sync();
    constexpr uint32_t idesc = ptx::idesc_bf16_f32(128, 64);
    const uint64_t a_tmpl = ptx::smem_desc(A, P1 * 16, PU1 * 16);
#pragma unroll 1
    for (int c = warp; c < chunks_of(W1); c += 2) {
      const uint64_t b_chunk = chunk_ready(G1 + c, 64 * 16);
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const int tap = 2 * c + h2;
        if (tap < 9) {
          const int di = tap / 3, dj = tap - 3 * di;
          const uint64_t a_tap = a_tmpl + (uint64_t)((uint32_t)(((di & 1) * 2 + (dj & 1)) * SUB1_BYTES + ((dj >> 1) * PU1 + (di >> 1)) * 16) >> 4);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            if (leader) ptx::mma_bf16_ss(acc_mine, a_tap + (uint32_t)((kk * 2 * P1 * 16) >> 4),
                                         b_chunk + (uint32_t)(((h2 * 8 + kk) * 64 * 32) >> 4), idesc, (c >= 2 || h2 || kk) ? 1u : 0u);
          }
        }
      }
      if (leader) ptx::mma_commit(bar(ring_empty(ring_slot(G1 + c))));
    }
    if (leader) ptx::mma_commit(bar(SB_TF));
    __syncwarp();
  }
  // stage 2's shortcut, front channel pad (reference :158-163): channels 64.. (= rank 1's) take the 2x2 average pool of x4.  The
  // four input pixels of output pixel (px, py) are entry [px][py] of the four parity sub-planes and x is exactly
  // recoverable from [elu(x) | elu(-x)], so the pool comes from shared memory while stage 1's MMAs read the same planes
  // (the scattered global loads it replaces cost the stage-2 epilogue 6 k clocks)
  float pool[NG][8];
  if (rank == 1) {
    const uint32_t e0 = A + (uint32_t)(px * PU1 + py) * 16u;
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const int pl = part * NG + g;
#pragma unroll
      for (int i = 0; i < 8; ++i) pool[g][i] = 0.f;
#pragma unroll
      for (int sp = 0; sp < 4; ++sp) {
        bf16x8 pos, neg;
        pos.v = lds128(e0 + (uint32_t)(sp * SUB1_BYTES + pl * P1 * 16));
        neg.v = lds128(e0 + (uint32_t)(sp * SUB1_BYTES + (8 + pl) * P1 * 16));
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

  // ---- epilogue of a plain conv (this CTA's 64 channels): acc + bias -> bf16 -> [elu | elu(-.)] planes, both CTAs
  auto epilogue_plain = [&](int boff) {
#pragma unroll 1
    for (int g = 0; g < NG; ++g) {
      const int cl = part * (8 * NG) + g * 8;                 // local channel
      const int cg = (int)rank * CL + cl;               // channel of the full tensor
      float o[8], bv[8];
      lds_f8(sb + 4u * (uint32_t)(boff + cl), bv);
      {
        float o2[8];
        ptx::tmem_ld8x2(lane_base + (uint32_t)cl, lane_base + (uint32_t)(128 + cl), o, o2);       // the two K halves
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] += o2[i];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] += bv[i];
      uint4 o0, o1;
      concat_elu8(pack8(o), o0, o1);
      const uint32_t d0 = (uint32_t)((cg >> 3) * PPAD * 16) + pp16;
      put_plane(d0, o0);
      put_plane(d0 + (uint32_t)(16 * PPAD * 16), o1);
    }
  };
  // the zero border of the stage-2..5 planes (the region held stage 1's sub-planes until now); local only: the peer
  // zeroes its own, and remote writes touch interior pixels only
  for (int idx = tid; idx < P * 32; idx += WT) {
    const int pl = idx / P, pp = idx - pl * P;
    const int pv = pp / PU, pu = pp - pv * PU;
    if (pu == 0 || pu == PU - 1 || pv == 0 || pv == PV - 1) sts128(A + (uint32_t)(pl * PPAD + pp) * 16u, make_uint4(0u, 0u, 0u, 0u));
  }
  peer_planes_free();
  epilogue_plain(B1);
  planes_exchanged();
  stamp();

  // ---- MMAs of a stride-1 3x3 conv on the 256-channel concat planes (stages 2, 3, 4), this CTA's N columns
  auto issue_s1 = [&](auto nconst, int g0) {
    constexpr int N = decltype(nconst)::value;
    constexpr int SPC = CHUNK / (N * 32);            // k-steps per chunk
    constexpr int CPT = 16 / SPC;                    // chunks per tap (16 k-steps)
    constexpr uint32_t idesc = ptx::idesc_bf16_f32(128, N);
    ptx::tc_fence_after_sync();
    const uint64_t a_tmpl = ptx::smem_desc(A, PPAD * 16, PU * 16);
#pragma unroll 1
    for (int c = warp; c < 144 / SPC; c += 2) {
      const uint64_t b_chunk = chunk_ready(g0 + c, N * 16);
      const int tap = c / CPT, kk0 = (c - tap * CPT) * SPC;
      const int di = tap / 3, dj = tap - 3 * di;
      const uint64_t a_tap = a_tmpl + (uint64_t)((uint32_t)((dj * PU + di) * 16 + kk0 * 2 * PPAD * 16) >> 4);
#pragma unroll
      for (int i = 0; i < SPC; ++i) {
        if (leader) ptx::mma_bf16_ss(acc_mine, a_tap + (uint32_t)((i * 2 * PPAD * 16) >> 4), b_chunk + (uint32_t)((i * N * 32) >> 4), idesc,
                                     (c >= 2 || i) ? 1u : 0u);
      }
      if (leader) ptx::mma_commit(bar(ring_empty(ring_slot(g0 + c))));
    }
    if (leader) ptx::mma_commit(bar(SB_TF));
    __syncwarp();
  };
  using N64 = std::integral_constant<int, 64>;
  using N128 = std::integral_constant<int, 128>;

  // ---- stage 2: conv_2 of the down block, gate, shortcut = 2x2 average pool of x4 in the LAST 64 channels (= rank 1's)
  if (warp < 2) issue_s1(N128{}, G2);
  stage_done();
  stamp();
  peer_planes_free();
  {
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const int cl = part * (8 * NG) + g * 8;
      const int cg = (int)rank * CL + cl;
      float a[8], gt[8], bu[8], bv[8], o[8];
      lds_f8(sb + 4u * (uint32_t)(B2 + cl), bu);
      lds_f8(sb + 4u * (uint32_t)(B2 + CL + cl), bv);
      ptx::tmem_ld8x2(lane_base + (uint32_t)cl, lane_base + (uint32_t)(CL + cl), a, gt);
      {
        float a2[8], gt2[8];
        ptx::tmem_ld8x2(lane_base + (uint32_t)(128 + cl), lane_base + (uint32_t)(128 + CL + cl), a2, gt2);      // the second K half
#pragma unroll
        for (int i = 0; i < 8; ++i) { a[i] += a2[i]; gt[i] += gt2[i]; }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float th = tanh_approx(fmaf(gt[i], 0.5f, bv[i]));
        o[i] = (a[i] + bu[i]) * fmaf(th, 0.5f, 0.5f);
      }
      if (rank == 1) {                 // cg >= C5 - C4: the pooled x4, channel cg - 64 = cl
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] += pool[g][i];
      }
      const uint4 raw = pack8(o);
      sts128(r1_row + (uint32_t)(((cl >> 3) ^ (r & 7)) * 16), raw);           // kept (own channels) for stage 4's shortcut
      uint4 o0, o1;
      concat_elu8(raw, o0, o1);
      const uint32_t d0 = (uint32_t)((cg >> 3) * PPAD * 16) + pp16;
      put_plane(d0, o0);
      put_plane(d0 + (uint32_t)(16 * PPAD * 16), o1);
    }
  }
  planes_exchanged();
  stamp();

  // ---- stage 3: conv_1 of resnet_5_0
  if (warp < 2) issue_s1(N64{}, G3);
  stage_done();
  stamp();
  peer_planes_free();
  epilogue_plain(B3);
  planes_exchanged();
  stamp();

  // ---- stage 4: conv_2 of resnet_5_0, gate, shortcut = the block input kept in shared memory; the result feeds the
  // transposed conv RAW (no nonlinearity, reference :213): 16 planes
  if (warp < 2) issue_s1(N128{}, G4);
  stage_done();
  stamp();
  peer_planes_free();
#pragma unroll 1
  for (int g = 0; g < NG; ++g) {
    const int cl = part * (8 * NG) + g * 8;
    const int cg = (int)rank * CL + cl;
    float a[8], gt[8], bu[8], bv[8], o[8];
    lds_f8(sb + 4u * (uint32_t)(B4 + cl), bu);
    lds_f8(sb + 4u * (uint32_t)(B4 + CL + cl), bv);
    ptx::tmem_ld8x2(lane_base + (uint32_t)cl, lane_base + (uint32_t)(CL + cl), a, gt);
    {
      float a2[8], gt2[8];
      ptx::tmem_ld8x2(lane_base + (uint32_t)(128 + cl), lane_base + (uint32_t)(128 + CL + cl), a2, gt2);        // the second K half
#pragma unroll
      for (int i = 0; i < 8; ++i) { a[i] += a2[i]; gt[i] += gt2[i]; }
    }
    bf16x8 rv;
    rv.v = lds128(r1_row + (uint32_t)(((cl >> 3) ^ (r & 7)) * 16));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float th = tanh_approx(fmaf(gt[i], 0.5f, bv[i]));
      // product and shortcut add rounded separately, as in the layer-by-layer kernels
      o[i] = __fadd_rn(__fmul_rn(a[i] + bu[i], fmaf(th, 0.5f, 0.5f)), __bfloat162float(rv.h[i]));
    }
    put_plane((uint32_t)((cg >> 3) * PPAD * 16) + pp16, pack8(o));
  }
  planes_exchanged();
  stamp();

  // ---- stage 5: transposed conv as four output phases, K = 128 per tap, this CTA's N = 32 per phase;
  // a 32 KB chunk = 32 k-steps = four taps
  if (warp == 0) {
    ptx::tc_fence_after_sync();
    constexpr uint32_t idesc = ptx::idesc_bf16_f32(128, 32);
    const uint64_t a_tmpl = ptx::smem_desc(A, PPAD * 16, PU * 16);
#pragma unroll 1
    for (int c = 0; c < chunks_of(W5); ++c) {
      const uint64_t b_chunk = chunk_ready(G5 + c, 32 * 16);
#pragma unroll
      for (int h4 = 0; h4 < 4; ++h4) {
        const int tap = 4 * c + h4;
        if (tap < 9) {
          const int di = tap / 3, dj = tap - 3 * di;
          // out[2m + p] takes tap d = p from in[m] and, for p = 0, tap d = 2 from in[m - 1]
          const uint64_t a_tap = a_tmpl + (uint64_t)((uint32_t)((((dj == 2) ? 0 : 1) * PU + ((di == 2) ? 0 : 1)) * 16) >> 4);
          const uint32_t d_tmem = tmem_base + (uint32_t)((((di == 1) ? 2 : 0) + ((dj == 1) ? 1 : 0)) * 32);
          const bool first_tap = tap == 0 || tap == 1 || tap == 3 || tap == 4;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            if (leader) ptx::mma_bf16_ss(d_tmem, a_tap + (uint32_t)((kk * 2 * PPAD * 16) >> 4),
                                         b_chunk + (uint32_t)(((h4 * 8 + kk) * 32 * 32) >> 4), idesc, (first_tap && kk == 0) ? 0u : 1u);
          }
        }
      }
      if (leader) ptx::mma_commit(bar(ring_empty(ring_slot(G5 + c))));
    }
    if (leader) ptx::mma_commit(bar(SB_TF));
    __syncwarp();
  }
  if (warp == 1) {               // the stage barrier counts both issuing warps; this stage is small enough for one
    if (leader) ptx::mma_commit(bar(SB_TF));
    __syncwarp();
  }
  stage_done();
  stamp();
  {
    __nv_bfloat16* yb = p.y + (size_t)b * IH * IW * C4;
#pragma unroll 1
    for (int ph = 0; ph < 4; ++ph) {
      const int oy = 2 * py + (ph >> 1), ox = 2 * px + (ph & 1);
      __nv_bfloat16* dst = yb + ((size_t)oy * IW + ox) * C4;
#pragma unroll
      for (int g = 0; g < NG5; ++g) {
        const int cl = part * (8 * NG5) + g * 8;        // 0..31
        const int cg = (int)rank * 32 + cl;             // output channel of up_conv_1
        float o[8], bv[8];
        lds_f8(sb + 4u * (uint32_t)(B5 + cl), bv);
        ptx::tmem_ld8(lane_base + (uint32_t)(ph * 32 + cl), o);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] += bv[i];
        if (ok) {
          const uint4 st = pack8(o);
          *reinterpret_cast<uint4*>(dst + cg) = st;
          if (p.yp) {     // planes form of the output (fn_level5_desc.y_planes): [B][16][IH][IW][8] = concat_elu(y)
            uint4 e0, e1;
            concat_elu8(st, e0, e1);
            __nv_bfloat16* d0 = p.yp + ((((size_t)b * 16 + (cg >> 3)) * IH + oy) * IW + ox) * 8;
            *reinterpret_cast<uint4*>(d0) = e0;
            *reinterpret_cast<uint4*>(d0 + (size_t)8 * IH * IW * 8) = e1;
          }
        }
      }
    }
  }
}

}  // namespace l5c
}  // namespace fn
