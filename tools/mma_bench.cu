// Microbenchmark: cycles per tcgen05.mma (M=128, K=16, bf16) as a function of N and of the
// shared-memory operand layout.  Answers "is the SWIZZLE_NONE tap-shift layout of conv_tc.cu slower
// to fetch than the canonical layouts?".  One CTA, one issuing thread, R back-to-back MMAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I steady-state-flow-with-neural-nets_b200/csrc \
//        -I include tools/mma_bench.cu -o tools/mma_bench
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "ptx.cuh"

using namespace fn;

__device__ __forceinline__ uint64_t desc_generic(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout_type) {
  uint64_t d = ptx::smem_desc(addr, lbo, sbo);
  d |= (uint64_t)layout_type << 61;
  return d;
}

struct Cfg {
  int N, R, ctas;
  uint32_t a_lbo, a_sbo, a_type;
  uint32_t b_lbo, b_sbo, b_type;
  uint32_t a_step, b_step;   // address advance per MMA (bytes), to mimic tap shifts / k-steps
  int J;                     // MMAs that share B per step
  int mode;                  // 0: rebuild descriptors per MMA under threadIdx==0; 1: base desc + 64-bit add under threadIdx==0;
                             // 2: whole warp runs the loop, MMA predicated by elect.sync; 3: as 2, unrolled x8
};

__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .b32 %%rx;\n\t.reg .pred %%px;\n\telect.sync %%rx|%%px, %1;\n\t@%%px mov.s32 %0, 1;\n\t}"
               : "+r"(pred)
               : "r"(0xFFFFFFFFu));
  return pred;
}

__global__ void __launch_bounds__(128) k_bench(Cfg c, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  // fill smem with small bf16 values (0x3c00 = 0.0078)
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) {
    ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), 512);
    ptx::tmem_relinquish();
  }
  if (threadIdx.x == 0) {
    ptx::mbar_init(ptx::smem_u32(&bar), 1);
    ptx::fence_mbar_init();
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = ptx::idesc_bf16_f32(128, c.N);
  const uint32_t a0 = (ptx::smem_u32(smem) + 1023u) & ~1023u, b0 = a0 + 96 * 1024;
  const uint32_t barp = ptx::smem_u32(&bar);
  if (c.mode <= 1) {
    if (threadIdx.x == 0) {
      const long long t0 = clock64();
      if (c.mode == 0) {
        for (int r = 0; r < c.R; ++r) {
          const uint64_t bd = desc_generic(b0 + (uint32_t)(r & 7) * c.b_step, c.b_lbo, c.b_sbo, c.b_type);
          for (int j = 0; j < c.J; ++j) {
            const uint64_t ad = desc_generic(a0 + (uint32_t)(r & 7) * c.a_step + (uint32_t)j * 128u, c.a_lbo, c.a_sbo, c.a_type);
            ptx::mma_bf16_ss(tmem + (uint32_t)(j * c.N) % 512u, ad, bd, idesc, 1u);
          }
        }
      } else {
        const uint64_t abase = desc_generic(a0, c.a_lbo, c.a_sbo, c.a_type);
        const uint64_t bbase = desc_generic(b0, c.b_lbo, c.b_sbo, c.b_type);
        const uint64_t astep = c.a_step >> 4, bstep = c.b_step >> 4;
        for (int r = 0; r < c.R; ++r) {
          const uint64_t bd = bbase + (uint64_t)(r & 7) * bstep;
          const uint64_t ad = abase + (uint64_t)(r & 7) * astep;
          for (int j = 0; j < c.J; ++j) ptx::mma_bf16_ss(tmem + (uint32_t)(j * c.N) % 512u, ad + 8u * j, bd, idesc, 1u);
        }
      }
      const long long t1 = clock64();
      ptx::mma_commit(barp);
      ptx::mbar_wait(barp, 0);
      const long long t2 = clock64();
      out[blockIdx.x * 2 + 0] = t1 - t0;
      out[blockIdx.x * 2 + 1] = t2 - t0;
    }
  } else if (warp == 1) {
    const uint32_t leader = elect_one_sync();
    const uint64_t abase = desc_generic(a0, c.a_lbo, c.a_sbo, c.a_type);
    const uint64_t bbase = desc_generic(b0, c.b_lbo, c.b_sbo, c.b_type);
    const uint64_t astep = c.a_step >> 4, bstep = c.b_step >> 4;
    const long long t0 = clock64();
    if (c.mode == 2) {
      for (int r = 0; r < c.R; ++r) {
        const uint64_t bd = bbase + (uint64_t)(r & 7) * bstep;
        const uint64_t ad = abase + (uint64_t)(r & 7) * astep;
        for (int j = 0; j < c.J; ++j)
          if (leader) ptx::mma_bf16_ss(tmem + (uint32_t)(j * c.N) % 512u, ad + 8u * j, bd, idesc, 1u);
      }
    } else {
      for (int r = 0; r < c.R; r += 8) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint64_t bd = bbase + (uint64_t)q * bstep;
          const uint64_t ad = abase + (uint64_t)q * astep;
          for (int j = 0; j < c.J; ++j)
            if (leader) ptx::mma_bf16_ss(tmem + (uint32_t)(j * c.N) % 512u, ad + 8u * j, bd, idesc, 1u);
        }
      }
    }
    __syncwarp();
    const long long t1 = clock64();
    if (leader) ptx::mma_commit(barp);
    __syncwarp();
    ptx::mbar_wait(barp, 0);
    const long long t2 = clock64();
    if (leader) {
      out[blockIdx.x * 2 + 0] = t1 - t0;
      out[blockIdx.x * 2 + 1] = t2 - t0;
    }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem, 512);
  }
}

// ---- cta_group::2: one tcgen05.mma over a CTA pair (M = 256: each SM computes its own 128 rows from its own shared
// memory, B is split N/2 per CTA), issued by the even CTA.  Question: does a pair instruction cost what a single-SM one
// costs (then the pair does twice the rows per instruction-slot) or twice that?
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) k_bench2(Cfg c, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    ptx::mbar_init(ptx::smem_u32(&bar), 1);
    ptx::fence_mbar_init();
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  ptx::tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = ptx::idesc_bf16_f32(256, c.N);
  const uint32_t a0 = (ptx::smem_u32(smem) + 1023u) & ~1023u, b0 = a0 + 96 * 1024;
  const uint32_t barp = ptx::smem_u32(&bar);
  if (rank == 0 && threadIdx.x == 0) {
    const uint64_t abase = desc_generic(a0, c.a_lbo, c.a_sbo, c.a_type);
    const uint64_t bbase = desc_generic(b0, c.b_lbo, c.b_sbo, c.b_type);
    const uint64_t astep = c.a_step >> 4, bstep = c.b_step >> 4;
    const long long t0 = clock64();
    for (int r = 0; r < c.R; r += 8) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint64_t bd = bbase + (uint64_t)q * bstep;
        const uint64_t ad = abase + (uint64_t)q * astep;
        for (int j = 0; j < c.J; ++j) {
          const uint32_t d = tmem + (uint32_t)(j * c.N) % 512u;
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad + 8u * j), "l"(bd), "r"(idesc), "r"(1u)
              : "memory");
        }
      }
    }
    const long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(barp) : "memory");
    ptx::mbar_wait(barp, 0);
    const long long t2 = clock64();
    out[(blockIdx.x / 2) * 2 + 0] = t1 - t0;
    out[(blockIdx.x / 2) * 2 + 1] = t2 - t0;
  }
  ptx::tc_fence_before_sync();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) {
    ptx::tc_fence_after_sync();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}


// ---- what is the fixed cost of a tcgen05.mma made of?  Straight-line issue (compile-time J, descriptors = base + immediates),
// R outer rounds of 8 k-steps x J MMAs; the slope between two R values removes the start / drain latency.
//   ACC  0: the J MMAs of a k-step go to J different accumulators (rotation)   1: all MMAs accumulate into ONE accumulator
//   BSAME 1: the J MMAs of a k-step share the B tile (as the conv kernels' taps do)   0: every MMA names a different B tile
// launched with 1 or 2 CTAs per SM (smem 96 KB, 256 TMEM columns each) to see whether independent CTAs overlap the cost.
template <int J, int ACC, int BSAME>
__global__ void __launch_bounds__(128) k_bench3(int N, int R, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) {
    ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), 256);
    ptx::tmem_relinquish();
  }
  if (threadIdx.x == 0) {
    ptx::mbar_init(ptx::smem_u32(&bar), 1);
    ptx::fence_mbar_init();
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = ptx::idesc_bf16_f32(128, N);
  const uint32_t a0 = ptx::smem_u32(smem), b0 = a0 + 32 * 1024;
  const uint32_t barp = ptx::smem_u32(&bar);
  if (warp == 1) {
    const uint32_t leader = elect_one_sync();
    const uint64_t abase = desc_generic(a0, 9808, 544, 0);
    const uint64_t bbase = desc_generic(b0, (uint32_t)N * 16, 128, 0);
    const uint32_t bstep = ((uint32_t)N * 32u) >> 4;      // one k-step of packed weights, in 16-byte units
    const int nb = 65536 / (N * 32);                      // B tiles that fit behind b0
    const long long t0 = clock64();
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
#pragma unroll
        for (int j = 0; j < J; ++j) {
          const int bi = BSAME ? q : (q * J + j) % nb;
          const uint32_t d = tmem + (ACC ? 0u : (uint32_t)(j * N) % 256u);
          if (leader) ptx::mma_bf16_ss(d, abase + (uint64_t)(q + 8 * j), bbase + (uint64_t)bi * bstep, idesc, 1u);
        }
      }
    }
    __syncwarp();
    if (leader) ptx::mma_commit(barp);
    __syncwarp();
    ptx::mbar_wait(barp, 0);
    const long long t2 = clock64();
    if (leader) out[blockIdx.x] = t2 - t0;
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem, 256);
  }
}

template <int J, int ACC, int BSAME>
static void run3(int N, long long* d_out) {
  if (!ACC && J * N > 256) return;
  cudaFuncSetAttribute(k_bench3<J, ACC, BSAME>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  double per[2];
  for (int two = 0; two < 2; ++two) {
    const int ctas = two ? 296 : 148;
    long long h[296];
    double t[2];
    for (int k = 0; k < 2; ++k) {
      const int R = k ? 64 : 16;
      cudaMemset(d_out, 0, 296 * sizeof(long long));
      k_bench3<J, ACC, BSAME><<<ctas, 128, 96 * 1024>>>(N, R, d_out);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return; }
      cudaMemcpy(h, d_out, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
      double s = 0;
      for (int i = 0; i < ctas; ++i) s += (double)h[i];
      t[k] = s / ctas;
    }
    per[two] = (t[1] - t[0]) / (48.0 * 8 * J);
  }
  printf("N=%3d J=%2d %s %s   1 CTA/SM: %6.1f clk/MMA   2 CTAs/SM: %6.1f clk/MMA per CTA = %6.1f per SM\n", N, J,
         ACC ? "one accumulator " : "rotating accs   ", BSAME ? "B shared by J" : "B differs     ", per[0], per[1], per[1] / 2);
}

int main() {
  long long* d3;
  cudaMalloc(&d3, 296 * sizeof(long long));
  if (getenv("MMA_BENCH3")) {
    for (int N : {16, 64, 128}) {
      run3<1, 0, 1>(N, d3); run3<2, 0, 1>(N, d3); run3<4, 0, 1>(N, d3); run3<8, 0, 1>(N, d3);
      run3<2, 0, 0>(N, d3); run3<4, 0, 0>(N, d3); run3<8, 0, 0>(N, d3);
      run3<4, 1, 1>(N, d3); run3<4, 1, 0>(N, d3);
    }
    run3<1, 0, 1>(256, d3); run3<4, 1, 1>(256, d3);
    return 0;
  }

  long long* d_out;
  cudaMalloc(&d_out, 148 * 2 * sizeof(long long));
  cudaFuncSetAttribute(k_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Named { const char* name; Cfg c; };
  const int R = 128;
  Named cases[] = {
      // name                                   N   R  ctas  a_lbo a_sbo at  b_lbo  b_sbo bt a_step b_step J
      {"noswz tapshift PU10 N16 J1",           {16, R, 1, 2880, 160, 0, 16 * 16, 128, 0, 16, 512, 1, 0}},
      {"noswz tapshift PU34 N16 J4",           {16, R, 1, 9808, 544, 0, 16 * 16, 128, 0, 16, 512, 4, 0}},
      {"noswz tapshift PU10 N64 J1",           {64, R, 1, 2896, 160, 0, 64 * 16, 128, 0, 16, 2048, 1, 0}},
      {"noswz tapshift PU10 N128 J1",          {128, R, 1, 2896, 160, 0, 128 * 16, 128, 0, 16, 4096, 1, 0}},
      {"noswz tapshift PU10 N256 J1",          {256, R, 1, 2896, 160, 0, 256 * 16, 128, 0, 16, 8192, 1, 0}},
      {"noswz canonical     N16 J1",           {16, R, 1, 2048, 128, 0, 16 * 16, 128, 0, 4096, 512, 1, 0}},
      {"noswz canonical     N64 J1",           {64, R, 1, 2048, 128, 0, 64 * 16, 128, 0, 4096, 2048, 1, 0}},
      {"noswz canonical     N256 J1",          {256, R, 1, 2048, 128, 0, 256 * 16, 128, 0, 4096, 8192, 1, 0}},
      {"swz128 canonical    N16 J1",           {16, R, 1, 16, 1024, 2, 16, 1024, 2, 32, 32, 1, 0}},
      {"swz128 canonical    N64 J1",           {64, R, 1, 16, 1024, 2, 16, 1024, 2, 32, 32, 1, 0}},
      {"swz128 canonical    N256 J1",          {256, R, 1, 16, 1024, 2, 16, 1024, 2, 32, 32, 1, 0}},
      {"swz128 A / noswz B  N16 J1",           {16, R, 1, 16, 1024, 2, 16 * 16, 128, 0, 32, 512, 1, 0}},
      {"noswz tapshift PU34 N16 J4 x148 CTAs", {16, R, 148, 9808, 544, 0, 16 * 16, 128, 0, 16, 512, 4, 0}},
      {"noswz tapshift PU10 N256 x148 CTAs",   {256, R, 148, 2896, 160, 0, 256 * 16, 128, 0, 16, 8192, 1, 0}},
      {"mode1 base+add   PU10 N16 J1",          {16, R, 1, 2880, 160, 0, 16 * 16, 128, 0, 16, 512, 1, 1}},
      {"mode1 base+add   PU34 N16 J4",          {16, R, 1, 9808, 544, 0, 16 * 16, 128, 0, 16, 512, 4, 1}},
      {"mode2 elect      PU10 N16 J1",          {16, R, 1, 2880, 160, 0, 16 * 16, 128, 0, 16, 512, 1, 2}},
      {"mode2 elect      PU34 N16 J4",          {16, R, 1, 9808, 544, 0, 16 * 16, 128, 0, 16, 512, 4, 2}},
      {"mode3 elect unr8 PU10 N16 J1",          {16, R, 1, 2880, 160, 0, 16 * 16, 128, 0, 16, 512, 1, 3}},
      {"mode3 elect unr8 PU34 N16 J4",          {16, R, 1, 9808, 544, 0, 16 * 16, 128, 0, 16, 512, 4, 3}},
      {"mode3 elect unr8 PU10 N32 J1",          {32, R, 1, 2880, 160, 0, 32 * 16, 128, 0, 16, 1024, 1, 3}},
      {"mode3 elect unr8 PU10 N64 J1",          {64, R, 1, 2896, 160, 0, 64 * 16, 128, 0, 16, 2048, 1, 3}},
      {"mode3 elect unr8 PU10 N128 J1",         {128, R, 1, 2896, 160, 0, 128 * 16, 128, 0, 16, 4096, 1, 3}},
      {"mode3 elect unr8 PU10 N256 J1",         {256, R, 1, 2896, 160, 0, 256 * 16, 128, 0, 16, 8192, 1, 3}},
      {"mode3 elect unr8 canon N16 J1",         {16, R, 1, 2048, 128, 0, 16 * 16, 128, 0, 4096, 512, 1, 3}},
      {"mode3 elect unr8 canon N256 J1",        {256, R, 1, 2048, 128, 0, 256 * 16, 128, 0, 4096, 8192, 1, 3}},
      {"mode3 elect unr8 swz128 N16 J1",        {16, R, 1, 16, 1024, 2, 16, 1024, 2, 32, 32, 1, 3}},
      {"mode3 elect unr8 swz128 N64 J1",        {64, R, 1, 16, 1024, 2, 16, 1024, 2, 32, 32, 1, 3}},
      {"mode3 elect unr8 swz128 N256 J1",       {256, R, 1, 16, 1024, 2, 16, 1024, 2, 32, 32, 1, 3}},
      {"mode3 PU34 N16 J4 x148 CTAs",           {16, R, 148, 9808, 544, 0, 16 * 16, 128, 0, 16, 512, 4, 3}},
  };
  // ---- accumulators in rotation (J) and the CTA-pair instruction, tap-shifted no-swizzle operands as in the conv kernels
  {
    cudaFuncSetAttribute(k_bench2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int Ns[] = {16, 32, 64, 128, 256};
    const int Js[] = {1, 2, 4, 8, 16};
    for (int N : Ns)
      for (int J : Js) {
        if (J * N > 512 && J > 1) continue;
        Cfg c = {N, R, 1, 9808, 544, 0, (uint32_t)N * 16, 128, 0, 16, (uint32_t)N * 32, J, 3};
        long long h1[2] = {0, 0}, h2[2] = {0, 0};
        cudaMemset(d_out, 0, 148 * 2 * sizeof(long long));
        k_bench<<<1, 128, 200 * 1024>>>(c, d_out);
        cudaError_t e1 = cudaDeviceSynchronize();
        cudaMemcpy(h1, d_out, sizeof(h1), cudaMemcpyDeviceToHost);
        cudaMemset(d_out, 0, 148 * 2 * sizeof(long long));
        k_bench2<<<2, 128, 200 * 1024>>>(c, d_out);
        cudaError_t e2 = cudaDeviceSynchronize();
        cudaMemcpy(h2, d_out, sizeof(h2), cudaMemcpyDeviceToHost);
        // all 74 pairs at once
        long long h3[2] = {0, 0};
        cudaMemset(d_out, 0, 148 * 2 * sizeof(long long));
        k_bench2<<<148, 128, 200 * 1024>>>(c, d_out);
        cudaError_t e3 = cudaDeviceSynchronize();
        cudaMemcpy(h3, d_out, sizeof(h3), cudaMemcpyDeviceToHost);
        const int mmas = R * J;
        printf("N=%3d J=%2d   1-SM M=128: %6.1f clk/MMA   CTA pair M=256: %6.1f clk/MMA  (74 pairs: %6.1f)   %s %s %s\n", N, J,
               (double)h1[1] / mmas, (double)h2[1] / mmas, (double)h3[1] / mmas, e1 == cudaSuccess ? "" : cudaGetErrorString(e1),
               e2 == cudaSuccess ? "" : cudaGetErrorString(e2), e3 == cudaSuccess ? "" : cudaGetErrorString(e3));
        if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) return 1;
      }
  }
  for (auto& nc : cases) {
    cudaMemset(d_out, 0, 148 * 2 * sizeof(long long));
    k_bench<<<nc.c.ctas, 128, 200 * 1024>>>(nc.c, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    const int mmas = nc.c.R * nc.c.J;
    printf("%-40s issue %7.1f clk/MMA   issue+complete %7.1f clk/MMA  (floor N/2 = %d)  %s\n", nc.name,
           (double)h[0] / mmas, (double)h[1] / mmas, nc.c.N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
    if (e != cudaSuccess) break;
  }
  return 0;
}
