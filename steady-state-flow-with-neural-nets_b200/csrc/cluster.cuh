// Thread-block-cluster plumbing shared by the fused level kernels: distributed-shared-memory stores, remote mbarrier
// arrives, cluster-scope waits, and the small epilogue helpers they all use.
#pragma once
#include "common.cuh"
#include "ptx.cuh"
#include "tma.cuh"

namespace fn {
namespace cl {
using tma::lds128;

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
// bounded (about a second), like ptx::mbar_wait
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
// x from one 8-channel group of [elu(x) | elu(-x)]: x > 0: elu(x) = x; x <= 0: elu(-x) = -x (exact in bf16)
__device__ __forceinline__ uint4 raw_from_concat_elu(const uint4& pos_v, const uint4& neg_v) {
  bf16x8 pos, neg, out;
  pos.v = pos_v;
  neg.v = neg_v;
#pragma unroll
  for (int i = 0; i < 8; ++i) out.h[i] = __bfloat162float(pos.h[i]) > 0.f ? pos.h[i] : __hneg(neg.h[i]);
  return out.v;
}

}  // namespace cl
}  // namespace fn
