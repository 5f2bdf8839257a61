// The 8x16 level of flow_net in one launch, two CTAs per image: the kernel around the stages of conv_l5c_body.cuh (where
// the design is described); the same stages also run inside the merged levels-4+5 kernel of conv_l4.cu.
#include "conv_l5c_body.cuh"

namespace fn {
namespace l5c {

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS) k_level5c(const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = ptx::smem_u32(smem);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + 256);
  float* sbias = reinterpret_cast<float*>(smem + 512);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int tid = threadIdx.x;
  const uint32_t rank = cluster_rank();

  if (warp == NWORK) {
    ptx::tmem_alloc(sbase + 256, 256u);
    ptx::tmem_relinquish();
    if (lane == 0) {
      init_barriers(sbase, NWORK);
      ptx::fence_mbar_init();
    }
    __syncwarp();
  }
  load_biases(sbias, p.bias, rank, tid, THREADS);
  ptx::tc_fence_before_sync();
  cluster_sync_all();                 // both CTAs' barriers exist before anything arrives on them
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  bool ok = true;
  Ctx c;
  c.sbase = sbase; c.bar0 = sbase; c.sb = sbase + 512u; c.tmem_base = tmem_base;
  c.warp = warp; c.lane = lane; c.tid = tid; c.b = (int)blockIdx.x >> 1; c.rank = rank;
  c.x = p.x; c.y = p.y; c.yp = p.yp; c.wstream = p.wstream; c.status = p.status;
  c.stamps = (p.timing != nullptr && (int)blockIdx.x < p.timing_ctas && tid == 0) ? p.timing + (size_t)blockIdx.x * 16 : nullptr;
  c.nstamp = 0;
  if (c.stamps) c.stamps[c.nstamp++] = clock64();

  if (warp == NWORK) producer(c, ok);
  else workers<1, false, NWORK>(c, ok);

  if (c.stamps) c.stamps[c.nstamp++] = clock64();
  // ---- neither CTA may leave while the other can still write into it or arrive on its barriers
  ptx::tc_fence_before_sync();
  cluster_sync_all();
  if (warp == NWORK) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, 256u);
  }
}

}  // namespace l5c

int* tc_status_word();   // conv_tc.cu
void tc_timing_buffer(long long** buf, int* ctas);   // conv_tc.cu: debug stamp buffer (fn_debug_set_timing)

int level5c_fwd(const fn_level5_desc& d, cudaStream_t stream) {
  using namespace l5c;
  Params p;
  p.x = (const __nv_bfloat16*)d.x;
  p.wstream = (const unsigned char*)d.w_stream;
  p.bias = d.bias;
  p.y = (__nv_bfloat16*)d.y;
  p.yp = (__nv_bfloat16*)d.y_planes;
  p.status = tc_status_word();
  if (p.status == nullptr) return (int)cudaErrorMemoryAllocation;
  p.B = d.B;
  tc_timing_buffer(&p.timing, &p.timing_ctas);
  p.timing_ctas /= 2;           // 16 stamps per CTA out of a buffer of 8 per registered CTA
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_level5c, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(level5c): %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  k_level5c<<<2 * d.B, THREADS, SMEM, stream>>>(p);
  count_launch();
  note_tcgen05(true);
  return check_launch("k_level5c");
}

long long level5c_weight_bytes() { return 2 * l5c::WTOTAL; }

}  // namespace fn
