// Small HBM-bound kernels around the convolutions: standalone nonlinearities (the public
// concat_elu()/set_nonlinearity() API, flow_architecture.py:14-29), dtype casts at the
// boundary, tf.nn.l2_loss (flow_net.py:39-42) and TF1 ApplyAdam (flow_net.py:44-46).
// All are grid-stride, 16-byte vectorised where the layout allows, grid = 148 SMs x 8.
#include "common.cuh"

namespace fn {

constexpr int EW_THREADS = 256;
static inline int ew_grid(long long work_items) {
  long long g = (work_items + EW_THREADS - 1) / EW_THREADS;
  const long long cap = 148LL * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// x: [n, C] bf16 -> y: [n, mult*C] bf16 ; one thread per 8-channel chunk (C % 8 == 0) or per element
__global__ void k_nonlin_vec(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n,
                             int C, int kind) {
  const int mult = pro_mult(kind);
  const int chunks = C / 8;
  const long long total = n * chunks;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / chunks;
    const int ch = (int)(i % chunks);
    bf16x8 in, o0, o1;
    in.v = *reinterpret_cast<const uint4*>(x + row * C + ch * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float p0, p1;
      pro_apply(kind, __bfloat162float(in.h[j]), p0, p1);
      o0.h[j] = __float2bfloat16(p0);
      o1.h[j] = __float2bfloat16(p1);
    }
    __nv_bfloat16* yr = y + row * (long long)(mult * C);
    *reinterpret_cast<uint4*>(yr + ch * 8) = o0.v;
    if (mult == 2) *reinterpret_cast<uint4*>(yr + C + ch * 8) = o1.v;
  }
}

__global__ void k_nonlin_scalar(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n,
                                int C, int kind) {
  const int mult = pro_mult(kind);
  const long long total = n * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / C;
    const int c = (int)(i % C);
    float p0, p1;
    pro_apply(kind, __bfloat162float(x[i]), p0, p1);
    y[row * (mult * C) + c] = __float2bfloat16(p0);
    if (mult == 2) y[row * (mult * C) + C + c] = __float2bfloat16(p1);
  }
}

int nonlinearity(const void* x, void* y, long long n, int C, int kind, cudaStream_t s) {
  if (n == 0) return 0;
  if ((C & 7) == 0 && aligned16(x) && aligned16(y)) {
    k_nonlin_vec<<<ew_grid(n * (C / 8)), EW_THREADS, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, n, C, kind);
  } else {
    k_nonlin_scalar<<<ew_grid(n * C), EW_THREADS, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, n, C, kind);
  }
  count_launch();
  return check_launch("k_nonlin");
}

template <typename TI, typename TO>
__device__ __forceinline__ TO cvt(TI v);
template <> __device__ __forceinline__ __nv_bfloat16 cvt<float, __nv_bfloat16>(float v) { return __float2bfloat16(v); }
template <> __device__ __forceinline__ __nv_bfloat16 cvt<uint8_t, __nv_bfloat16>(uint8_t v) { return __float2bfloat16((float)v); }
template <> __device__ __forceinline__ float cvt<__nv_bfloat16, float>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename TI, typename TO>
__global__ void k_cast(const TI* __restrict__ x, TO* __restrict__ y, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = cvt<TI, TO>(x[i]);
}

template <typename TI, typename TO>
int cast(const TI* x, TO* y, long long n, cudaStream_t s) {
  if (n == 0) return 0;
  k_cast<TI, TO><<<ew_grid(n), EW_THREADS, 0, s>>>(x, y, n);
  count_launch();
  return check_launch("k_cast");
}
template int cast<float, __nv_bfloat16>(const float*, __nv_bfloat16*, long long, cudaStream_t);
template int cast<uint8_t, __nv_bfloat16>(const uint8_t*, __nv_bfloat16*, long long, cudaStream_t);
template int cast<__nv_bfloat16, float>(const __nv_bfloat16*, float*, long long, cudaStream_t);

// loss += sum((p-t)^2)/2 ; grad = p - t.  Block tree-reduce then one atomicAdd per block.
__global__ void k_l2_loss(const float* __restrict__ p, const float* __restrict__ t, float* loss, float* grad,
                          long long n) {
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = p[i] - t[i];
    if (grad) grad[i] = d;
    s = fmaf(d, d, s);
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ float ws[EW_THREADS / 32];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < EW_THREADS / 32 ? ws[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) atomicAdd(loss, 0.5f * s);
  }
}

int l2_loss(const float* p, const float* t, float* loss, float* grad, long long n, cudaStream_t s) {
  if (n == 0) return 0;
  k_l2_loss<<<ew_grid(n), EW_THREADS, 0, s>>>(p, t, loss, grad, n);
  count_launch();
  return check_launch("k_l2_loss");
}

// TF1 ApplyAdam: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m,v EMA; theta -= lr_t*m/(sqrt(v)+eps)
__global__ void k_adam(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ m,
                       float* __restrict__ v, long long n, float lr_t, float b1, float omb1, float b2, float omb2,
                       float eps) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float g = grad[i];
    const float mi = fmaf(b1, m[i], omb1 * g);
    const float vi = fmaf(b2, v[i], omb2 * g * g);
    m[i] = mi;
    v[i] = vi;
    param[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

int adam_step(float* param, const float* grad, float* m, float* v, long long n, long long step, double lr, double b1,
              double b2, double eps, cudaStream_t s) {
  if (n == 0) return 0;
  // coefficients are formed in double on the host so (1-beta) is exact to fp32 rounding
  const double lr_t = lr * sqrt(1.0 - pow(b2, (double)step)) / (1.0 - pow(b1, (double)step));
  k_adam<<<ew_grid(n), EW_THREADS, 0, s>>>(param, grad, m, v, n, (float)lr_t, (float)b1, (float)(1.0 - b1), (float)b2,
                                           (float)(1.0 - b2), (float)eps);
  count_launch();
  return check_launch("k_adam");
}

}  // namespace fn
