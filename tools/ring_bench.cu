// Microbenchmark: how fast can ONE CTA per SM stream an L2-resident buffer into shared memory with cp.async.bulk through a
// ring of NST slots of CHUNK bytes, when a slot is refilled only after its consumer released it (the weight rings of
// conv_s1d / conv_l5c / conv_l4)?  Answers "is the ring bound by per-SM bandwidth or by round-trip latency x bytes in flight".
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I steady-state-flow-with-neural-nets_b200/csrc -I include tools/ring_bench.cu -o tools/ring_bench
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace fn;

// consumer holds every chunk for `hold` clocks (the MMA time of a chunk) before releasing it
__global__ void __launch_bounds__(64) k_ring(const unsigned char* src, int total_chunks, int nst, int chunk, int hold, int split, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bars[32];
  const uint32_t sbase = ptx::smem_u32(smem);
  auto full = [&](int s) { return ptx::smem_u32(&bars[s]); };
  auto empty = [&](int s) { return ptx::smem_u32(&bars[16 + s]); };
  if (threadIdx.x == 0) {
    for (int s = 0; s < nst; ++s) { ptx::mbar_init(full(s), 1); ptx::mbar_init(empty(s), 1); }
    ptx::fence_mbar_init();
  }
  __syncthreads();
  const long long t0 = clock64();
  if (threadIdx.x == 32) {           // producer
    const uint64_t pol = ptx::policy_evict_last();
    for (int gc = 0; gc < total_chunks; ++gc) {
      const int s = gc % nst;
      if (gc >= nst) ptx::mbar_wait(empty(s), (uint32_t)(((gc / nst) - 1) & 1));
      ptx::mbar_arrive_expect_tx(full(s), (uint32_t)chunk);
      for (int k = 0; k < split; ++k)
        ptx::bulk_g2s_hint(sbase + s * chunk + k * (chunk / split), src + (size_t)(gc % 64) * chunk + k * (chunk / split), (uint32_t)(chunk / split), full(s), pol);
    }
  } else if (threadIdx.x == 0) {     // consumer
    for (int gc = 0; gc < total_chunks; ++gc) {
      const int s = gc % nst;
      ptx::mbar_wait(full(s), (uint32_t)((gc / nst) & 1));
      const long long t = clock64();
      while (clock64() - t < hold) {}
      ptx::mbar_arrive(empty(s));
    }
    out[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  unsigned char* src;
  cudaMalloc(&src, 64 * 65536);
  cudaMemset(src, 1, 64 * 65536);
  long long* d_out;
  cudaMalloc(&d_out, 148 * sizeof(long long));
  cudaFuncSetAttribute(k_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int total_bytes = 4 << 20;
  for (int ctas : {128})
    for (int chunk : {4096, 16384, 32768})
      for (int nst : {1, 2, 4, 6}) {
        if (nst * chunk > 192 * 1024) continue;
        for (int hold : {0, 512, 1024})
        for (int split : {1, 4}) {
          const int n = total_bytes / chunk;
          k_ring<<<ctas, 64, 200 * 1024>>>(src, n, nst, chunk, hold, split, d_out);   // warm L2
          k_ring<<<ctas, 64, 200 * 1024>>>(src, n, nst, chunk, hold, split, d_out);
          if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
          long long h[148];
          cudaMemcpy(h, d_out, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
          double s = 0;
          for (int i = 0; i < ctas; ++i) s += (double)h[i];
          s /= ctas;
          printf("ctas %3d hold %4d copies/chunk %d chunk %5d x %2d slots (%3d KB ring): %6.1f B/clk per SM, %6.0f clk per chunk, implied round trip %5.0f clk\n", ctas, hold, split,
                 chunk, nst, nst * chunk / 1024, total_bytes / s, s / n, s / n * (nst));
        }
      }
  return 0;
}
