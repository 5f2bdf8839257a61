// D2Q9 BGK lattice-Boltzmann channel flow on the GPU: the simulation that labels the reference's training data
// (fluid_generator/generate_xiao_fluid_flow.cpp:147-188 boundary columns, :209-365 set-up; the time loop itself is
// MechSys', un-vendored -- restated in oracle/lbm_oracle.py, which this kernel follows operation by operation).
// replaces the MechSys dependency for generating labelled flows on the box.
//
// One fused kernel per time step, one thread per lattice cell, populations in structure-of-arrays order
// [B][9][ny][nx] so that all 9 loads of a warp are contiguous: Zou-He inlet / outlet columns -> macroscopic fields ->
// BGK collision (solids: full bounce-back) -> periodic push-streaming into the second lattice.  Pure HBM stream:
// 18 values per cell and step (72 B in fp32, 144 B in fp64), no reuse to exploit -- the roofline is HBM bandwidth.
#include "common.cuh"

namespace fn {
namespace lbm {

__constant__ int c_cx[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
__constant__ int c_cy[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};

template <typename T>
struct Args {
  const T* fin;
  T* fout;
  const uint8_t* solid;     // [B, ny, nx], 1 = solid
  float* vel;               // fields kernel: [B, ny, nx, 2]
  float* rho;               // fields kernel: [B, ny, nx] or null
  int B, nx, ny;
  T inv_tau, u_max, rho_out;
};

template <typename T>
__device__ __forceinline__ T inlet_velocity(int iy, int ny, T u_max) {      // :239-246
  const T L = (T)(ny - 2);
  const T yp = (T)iy - (T)1.5;
  return u_max * (T)4 / (L * L) * (L * yp - yp * yp);
}

// Zou-He columns (:150-157, :174-179), applied to every row as the reference does
template <typename T>
__device__ __forceinline__ void boundary_columns(T (&f)[9], int ix, int iy, const Args<T>& a) {
  if (ix == 0) {
    const T v = inlet_velocity<T>(iy, a.ny, a.u_max);
    const T rho = (f[0] + f[2] + f[4] + (T)2 * (f[3] + f[6] + f[7])) / ((T)1 - v);
    f[1] = f[3] + ((T)2 / (T)3) * rho * v;
    f[5] = f[7] + ((T)1 / (T)6) * rho * v - (T)0.5 * (f[2] - f[4]);
    f[8] = f[6] + ((T)1 / (T)6) * rho * v + (T)0.5 * (f[2] - f[4]);
  }
  if (ix == a.nx - 1) {
    const T vx = (T)-1 + (f[0] + f[2] + f[4] + (T)2 * (f[1] + f[5] + f[8])) / a.rho_out;
    f[3] = f[1] - ((T)2 / (T)3) * a.rho_out * vx;
    f[7] = f[5] - ((T)1 / (T)6) * a.rho_out * vx + (T)0.5 * (f[2] - f[4]);
    f[6] = f[8] - ((T)1 / (T)6) * a.rho_out * vx - (T)0.5 * (f[2] - f[4]);
  }
}

template <typename T>
__device__ __forceinline__ T weight(int k) { return k == 0 ? (T)4 / (T)9 : (k < 5 ? (T)1 / (T)9 : (T)1 / (T)36); }

// collision of one cell (after the boundary columns were closed): BGK for fluid, full bounce-back for solids
template <typename T>
__device__ __forceinline__ void collide(const T (&f)[9], bool is_solid, T inv_tau, T (&post)[9]) {
  if (is_solid) {
    const int opp[9] = {0, 3, 4, 1, 2, 7, 8, 5, 6};
#pragma unroll
    for (int k = 0; k < 9; ++k) post[k] = f[opp[k]];
    return;
  }
  T rho = 0, mx = 0, my = 0;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    rho += f[k];
    mx += (T)c_cx[k] * f[k];
    my += (T)c_cy[k] * f[k];
  }
  const T vx = mx / rho, vy = my / rho;
  const T vv = vx * vx + vy * vy;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const T cu = (T)c_cx[k] * vx + (T)c_cy[k] * vy;
    const T feq = weight<T>(k) * rho * ((T)1 + (T)3 * cu + (T)4.5 * cu * cu - (T)1.5 * vv);
    post[k] = f[k] - (f[k] - feq) * inv_tau;
  }
}

// grid = (ceil(nx / (128 * CELLS)), ny, B): no index division; a thread owns CELLS cells 128 columns apart, so every
// load / store of a warp stays contiguous and CELLS x 9 loads are in flight per thread (fp32 needs the second cell to
// keep enough bytes in flight; fp64 does not)
template <typename T, int CELLS>
__global__ void __launch_bounds__(128) k_lbm_step(const Args<T> a) {
  const int iy = blockIdx.y;
  const long long b = blockIdx.z;
  const long long plane = (long long)a.ny * a.nx;
  const long long row = b * plane + (long long)iy * a.nx;
  const T* __restrict__ fi = a.fin + b * 9 * plane + (long long)iy * a.nx;
  T f[CELLS][9];
  int ixs[CELLS];
  bool sol[CELLS];
#pragma unroll
  for (int c = 0; c < CELLS; ++c) {
    ixs[c] = (blockIdx.x * CELLS + c) * 128 + threadIdx.x;
    sol[c] = false;
    if (ixs[c] < a.nx) {
#pragma unroll
      for (int k = 0; k < 9; ++k) f[c][k] = __ldg(fi + k * plane + ixs[c]);
      sol[c] = a.solid[row + ixs[c]] != 0;
    }
  }
  T* __restrict__ fo = a.fout + b * 9 * plane;
  const int ym = iy == 0 ? a.ny - 1 : iy - 1, yp = iy == a.ny - 1 ? 0 : iy + 1;
  const int tys[3] = {ym, iy, yp};
#pragma unroll
  for (int c = 0; c < CELLS; ++c) {
    const int ix = ixs[c];
    if (ix >= a.nx) continue;
    boundary_columns<T>(f[c], ix, iy, a);
    T post[9];
    collide<T>(f[c], sol[c], a.inv_tau, post);
    const int xm = ix == 0 ? a.nx - 1 : ix - 1, xp = ix == a.nx - 1 ? 0 : ix + 1;
    const int txs[3] = {xm, ix, xp};
#pragma unroll
    for (int k = 0; k < 9; ++k) fo[k * plane + (long long)tys[c_cy[k] + 1] * a.nx + txs[c_cx[k] + 1]] = post[k];
  }
}

template <typename T>
__global__ void __launch_bounds__(256) k_lbm_init(T* f, int B, int nx, int ny, T rho0, T vx, T vy) {
  const long long plane = (long long)ny * nx;
  const long long total = (long long)B * 9 * plane;
  const T vv = vx * vx + vy * vy;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)((i / plane) % 9);
    const T cu = (T)c_cx[k] * vx + (T)c_cy[k] * vy;
    f[i] = weight<T>(k) * rho0 * ((T)1 + (T)3 * cu + (T)4.5 * cu * cu - (T)1.5 * vv);
  }
}

// macroscopic fields of the current populations (boundary columns closed first); solids report zero
template <typename T>
__global__ void __launch_bounds__(256) k_lbm_fields(const Args<T> a) {
  const long long plane = (long long)a.ny * a.nx;
  const long long cells = (long long)a.B * plane;
  for (long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x; cell < cells; cell += (long long)gridDim.x * blockDim.x) {
    const int ix = (int)(cell % a.nx);
    const int iy = (int)((cell / a.nx) % a.ny);
    const long long b = cell / plane;
    const T* fi = a.fin + b * 9 * plane + (long long)iy * a.nx + ix;
    T f[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) f[k] = fi[k * plane];
    boundary_columns<T>(f, ix, iy, a);
    T rho = 0, mx = 0, my = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      rho += f[k];
      mx += (T)c_cx[k] * f[k];
      my += (T)c_cy[k] * f[k];
    }
    const bool s = a.solid[cell] != 0;
    a.vel[cell * 2] = s ? 0.f : (float)(mx / rho);
    a.vel[cell * 2 + 1] = s ? 0.f : (float)(my / rho);
    if (a.rho) a.rho[cell] = s ? 0.f : (float)rho;
  }
}

static int grid_for(long long items) {
  long long g = (items + 255) / 256;
  const long long cap = 148LL * 16;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

template <typename T>
static int steps(const fn_lbm_desc& d, int nsteps, cudaStream_t s) {
  Args<T> a;
  memset(&a, 0, sizeof(a));
  a.solid = d.solid;
  a.B = d.B; a.nx = d.nx; a.ny = d.ny;
  a.inv_tau = (T)(1.0 / d.tau);
  a.u_max = (T)d.u_max;
  a.rho_out = (T)d.rho_out;
  T* fa = (T*)d.f_a;
  T* fb = (T*)d.f_b;
  for (int i = 0; i < nsteps; ++i) {
    a.fin = (i & 1) ? fb : fa;
    a.fout = (i & 1) ? fa : fb;
    constexpr int CELLS = sizeof(T) == 4 ? 3 : 1;
    k_lbm_step<T, CELLS><<<dim3((unsigned)((d.nx + 128 * CELLS - 1) / (128 * CELLS)), (unsigned)d.ny, (unsigned)d.B), 128, 0, s>>>(a);
  }
  count_launch(nsteps);
  return check_launch("k_lbm_step");
}

template <typename T>
static int fields(const fn_lbm_desc& d, float* vel, float* rho, cudaStream_t s) {
  Args<T> a;
  memset(&a, 0, sizeof(a));
  a.fin = (const T*)d.f_a;
  a.solid = d.solid;
  a.vel = vel;
  a.rho = rho;
  a.B = d.B; a.nx = d.nx; a.ny = d.ny;
  a.inv_tau = (T)(1.0 / d.tau);
  a.u_max = (T)d.u_max;
  a.rho_out = (T)d.rho_out;
  k_lbm_fields<T><<<grid_for((long long)d.B * d.nx * d.ny), 256, 0, s>>>(a);
  count_launch();
  return check_launch("k_lbm_fields");
}

static int validate(const fn_lbm_desc* d) {
  FN_REQUIRE(d != nullptr, FN_E_INVAL, "fn_lbm: null descriptor");
  FN_REQUIRE(d->B > 0 && d->B <= 65535 && d->nx >= 3 && d->ny >= 3 && d->ny <= 65535, FN_E_INVAL, "fn_lbm: bad lattice %d x %d x %d",
             d->B, d->ny, d->nx);
  FN_REQUIRE(d->dtype == 0 || d->dtype == 1, FN_E_INVAL, "fn_lbm: dtype %d (0 = fp32, 1 = fp64)", d->dtype);
  FN_REQUIRE(d->tau > 0.5, FN_E_INVAL, "fn_lbm: tau = %g must exceed 1/2", d->tau);
  FN_REQUIRE(d->solid && d->f_a && d->f_b, FN_E_INVAL, "fn_lbm: null solid / lattice buffers");
  return 0;
}

}  // namespace lbm
}  // namespace fn

using namespace fn;

extern "C" {

int fn_lbm_init(const fn_lbm_desc* d, double rho0, double vx0, double vy0, void* stream) {
  int rc = lbm::validate(d);
  if (rc) return rc;
  const long long total = (long long)d->B * 9 * d->nx * d->ny;
  if (d->dtype == 0)
    lbm::k_lbm_init<float><<<lbm::grid_for(total), 256, 0, (cudaStream_t)stream>>>((float*)d->f_a, d->B, d->nx, d->ny, (float)rho0,
                                                                                  (float)vx0, (float)vy0);
  else
    lbm::k_lbm_init<double><<<lbm::grid_for(total), 256, 0, (cudaStream_t)stream>>>((double*)d->f_a, d->B, d->nx, d->ny, rho0, vx0, vy0);
  count_launch();
  return check_launch("k_lbm_init");
}

int fn_lbm_steps(const fn_lbm_desc* d, int32_t nsteps, void* stream) {
  int rc = lbm::validate(d);
  if (rc) return rc;
  FN_REQUIRE(nsteps >= 0 && (nsteps & 1) == 0, FN_E_INVAL, "fn_lbm_steps: nsteps = %d must be even (the result stays in f_a)", nsteps);
  return d->dtype == 0 ? lbm::steps<float>(*d, nsteps, (cudaStream_t)stream) : lbm::steps<double>(*d, nsteps, (cudaStream_t)stream);
}

int fn_lbm_fields(const fn_lbm_desc* d, float* vel, float* rho, void* stream) {
  int rc = lbm::validate(d);
  if (rc) return rc;
  FN_REQUIRE(vel != nullptr, FN_E_INVAL, "fn_lbm_fields: null vel");
  return d->dtype == 0 ? lbm::fields<float>(*d, vel, rho, (cudaStream_t)stream) : lbm::fields<double>(*d, vel, rho, (cudaStream_t)stream);
}

}  // extern "C"
