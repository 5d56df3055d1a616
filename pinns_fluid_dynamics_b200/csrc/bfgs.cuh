// Quasi-Newton algebra of the BFGS round on the device (ns.minimize(pb, 'scipy', 'BFGS', epochs), cavity_steady.py:247):
// the dense float64 inverse Hessian (P x P: 42.6 MB for the 2307-parameter network) never leaves HBM, the host sees scalars.
//   trial point   x_t = x + alpha p, theta = float(x_t)                                  bfgs_trial_kernel
//   evaluation    g_t = double(grad), phi = sum_t c_t v_t, phi' = g_t . p, |g_t|_inf      bfgs_eval_kernel (one CTA: fixed-order sums)
//   accept        s = x_t - x, y = g_t - g, x = x_t, g = g_t, ys = y . s                  bfgs_accept_kernel
//   update        u = H y (warp per row), then in ONE pass over H:  H <- H - rho (s u^T + u s^T) + (rho^2 y.u + rho) s s^T
//                 and the next direction p = -H g of the updated rows                      bfgs_matvec_kernel, bfgs_update_direction_kernel
// rho = 1 / (y . s) (1000 when y . s == 0, like SciPy) is formed on the device: an iteration needs no host round trip
// between its kernels.  Every reduction has a fixed order (bit-reproducible iterates).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pinn {
namespace bfgs {

// scalars: [0] phi  [1] phi' = g_t . p  [2] |g_t|_inf  [3] y . s  [4] g . p of the new direction  [5] |p|_2
constexpr int kScalars = 8;

__global__ void bfgs_trial_kernel(const double* __restrict__ x, const double* __restrict__ p, double alpha, double* __restrict__ xt,
                                  float* __restrict__ theta, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const double v = fma(alpha, p[i], x[i]);
    xt[i] = v;
    theta[i] = (float)v;
  }
}

// the same with the step length read from device memory (the caller's CUDA graph replays with a new alpha every time)
__global__ void bfgs_trial_dev_kernel(const double* __restrict__ x, const double* __restrict__ p, const double* __restrict__ alpha,
                                      double* __restrict__ xt, float* __restrict__ theta, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const double v = fma(*alpha, p[i], x[i]);
    xt[i] = v;
    theta[i] = (float)v;
  }
}

__device__ __forceinline__ double block_sum(double v, double* sh) {      // fixed-order sum over a 1024-thread CTA
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  double s = 0.0;
  if (w == 0) {
    s = l < (int)(blockDim.x >> 5) ? sh[l] : 0.0;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (l == 0) sh[32] = s;
  }
  __syncthreads();
  return sh[32];
}
__device__ __forceinline__ double block_max(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  double s = 0.0;
  if (w == 0) {
    s = l < (int)(blockDim.x >> 5) ? sh[l] : 0.0;
    for (int o = 16; o > 0; o >>= 1) s = fmax(s, __shfl_xor_sync(0xffffffffu, s, o));
    if (l == 0) sh[32] = s;
  }
  __syncthreads();
  return sh[32];
}

// out: [P + T] of the loss step (gradient, then per-term sums in kernel order); coef[t] = weight / (normalization N_global)
// (0 for test terms), kind[t] = 1 for |mean| terms
__global__ void __launch_bounds__(1024) bfgs_eval_kernel(const float* __restrict__ out, const double* __restrict__ coef,
                                                         const int* __restrict__ kind, int n_terms, const double* __restrict__ p,
                                                         double* __restrict__ gt, double* __restrict__ scal, int64_t n) {
  __shared__ double sh[40];
  double dot = 0.0, mx = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double g = (double)out[i];
    gt[i] = g;
    dot = fma(g, p[i], dot);
    mx = fmax(mx, fabs(g));
  }
  dot = block_sum(dot, sh);
  mx = block_max(mx, sh);
  if (threadIdx.x == 0) {
    double phi = 0.0;
    for (int t = 0; t < n_terms; ++t) {
      const double v = (double)out[n + t];
      phi += coef[t] * (kind[t] ? fabs(v) : v);
    }
    scal[0] = phi;
    scal[1] = dot;
    scal[2] = mx;
  }
}

__global__ void __launch_bounds__(1024) bfgs_accept_kernel(double* __restrict__ x, double* __restrict__ g, const double* __restrict__ xt,
                                                           const double* __restrict__ gt, double* __restrict__ s, double* __restrict__ y,
                                                           double* __restrict__ scal, int64_t n) {
  __shared__ double sh[40];
  double ys = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double si = xt[i] - x[i], yi = gt[i] - g[i];
    s[i] = si;
    y[i] = yi;
    x[i] = xt[i];
    g[i] = gt[i];
    ys = fma(yi, si, ys);
  }
  ys = block_sum(ys, sh);
  if (threadIdx.x == 0) scal[3] = ys;
}

// u = H v, one warp per row (H row-major, symmetric)
__global__ void __launch_bounds__(256) bfgs_matvec_kernel(const double* __restrict__ H, const double* __restrict__ v, double* __restrict__ u,
                                                          double sign, int64_t n) {
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= n) return;
  const int lane = threadIdx.x & 31;
  const double* h = H + row * n;
  double acc = 0.0;
  for (int64_t j = lane; j < n; j += 32) acc = fma(h[j], v[j], acc);
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) u[row] = sign * acc;
}

// one pass over H: rank-2 update of row i and p_i = -(updated row) . g
__global__ void __launch_bounds__(256) bfgs_update_direction_kernel(double* __restrict__ H, const double* __restrict__ s,
                                                                    const double* __restrict__ y, const double* __restrict__ u,
                                                                    const double* __restrict__ g, double* __restrict__ p,
                                                                    const double* __restrict__ scal, int64_t n) {
  __shared__ double sh[40];
  // a = y . u, recomputed per CTA in a fixed order (n is a few thousand: cheaper than another launch)
  double a = 0.0;
  for (int64_t j = threadIdx.x; j < n; j += blockDim.x) a = fma(y[j], u[j], a);
  a = block_sum(a, sh);
  const double ys = scal[3];
  const double rho = ys == 0.0 ? 1000.0 : 1.0 / ys;
  const double c = rho * rho * a + rho;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= n) return;
  const int lane = threadIdx.x & 31;
  double* h = H + row * n;
  const double si = s[row], ui = u[row];
  double acc = 0.0;
  for (int64_t j = lane; j < n; j += 32) {
    const double hv = h[j] - rho * (si * u[j] + ui * s[j]) + c * si * s[j];
    h[j] = hv;
    acc = fma(hv, g[j], acc);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) p[row] = -acc;
}

// scal[4] = g . p, scal[5] = |p|_2
__global__ void __launch_bounds__(1024) bfgs_slope_kernel(const double* __restrict__ g, const double* __restrict__ p, double* __restrict__ scal,
                                                          int64_t n) {
  __shared__ double sh[40];
  double dot = 0.0, pp = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    dot = fma(g[i], p[i], dot);
    pp = fma(p[i], p[i], pp);
  }
  dot = block_sum(dot, sh);
  pp = block_sum(pp, sh);
  if (threadIdx.x == 0) {
    scal[4] = dot;
    scal[5] = sqrt(pp);
  }
}

__global__ void bfgs_identity_kernel(double* __restrict__ H, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n * n) H[i] = (i / n == i % n) ? 1.0 : 0.0;
}

}  // namespace bfgs
}  // namespace pinn
