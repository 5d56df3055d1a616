// FP32 FMA-pipe peak probe for B200 (sm_100a).
//
// MEASURED_PEAKS.json (driver-written) holds only the HBM copy bandwidth and the
// bf16 cuBLAS peak.  The PINN loss step for widths <= 32 is bounded by the FP32
// FMA pipe, so the roofline denominator for it has to be measured here: this
// program times long chains of independent FFMA (3-register form) and FFMA2
// (packed fma.rn.f32x2, new on sm_100) and prints one JSON line.  bench.py reads
// the result from profiles/fp32_peak_*.json (committed) and records which figure
// it used.
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/fp32_peak tools/fp32_peak.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int NACC = 64;

__global__ void __launch_bounds__(256) ffma_chain(float* out, float a, float b, int iters) {
  float acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = (float)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// GEMM-like pattern: acc[r][c] += x[r] * w[c], 8x8 register tile, operands refreshed from
// registers each step (no memory) -- the shape of the jet kernel's inner loop.
__global__ void __launch_bounds__(256) ffma_tile(float* out, const float* in, int iters) {
  float acc[8][8];
  float x[8], w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = in[threadIdx.x + i]; w[i] = in[threadIdx.x + 8 + i]; }
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[r][c] = fmaf(x[r], w[c], acc[r][c]);
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] += 1e-9f; }
  }
  float s = 0.f;
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < 8; ++c) s += acc[r][c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) ffma2_chain(float* out, float a, float b, int iters) {
  float2 acc[NACC / 2];
#pragma unroll
  for (int i = 0; i < NACC / 2; ++i) acc[i] = make_float2((float)(threadIdx.x + i), (float)i);
  const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC / 2; ++i) acc[i] = __ffma2_rn(acc[i], a2, b2);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC / 2; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// packed GEMM-like pattern: acc[r][c2] (float2 over 2 columns) += {x[r],x[r]} * w2[c2]
__global__ void __launch_bounds__(256) ffma2_tile(float* out, const float* in, int iters) {
  float2 acc[8][4];
  float2 x[8], w[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { float v = in[threadIdx.x + i]; x[i] = make_float2(v, v); }
#pragma unroll
  for (int i = 0; i < 4; ++i) w[i] = make_float2(in[threadIdx.x + 8 + 2 * i], in[threadIdx.x + 9 + 2 * i]);
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = make_float2(0.f, 0.f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = __ffma2_rn(x[r], w[c], acc[r][c]);
#pragma unroll
    for (int i = 0; i < 4; ++i) { w[i].x += 1e-9f; }
  }
  float s = 0.f;
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) s += acc[r][c].x + acc[r][c].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_best(F launch, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int i = 0; i < reps; ++i) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    best = std::min(best, (double)ms);
  }
  return best;
}

int main(int argc, char** argv) {
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
  const int sms = p.multiProcessorCount;
  const int blocks = sms * 8, threads = 256, iters = 4096;
  float *out, *in; CK(cudaMalloc(&out, sizeof(float) * blocks * threads));
  CK(cudaMalloc(&in, sizeof(float) * 1024)); CK(cudaMemset(in, 0, sizeof(float) * 1024));
  const double fl_chain = 2.0 * NACC * (double)iters * blocks * threads;
  const double fl_tile = 2.0 * 64 * (double)iters * blocks * threads;
  double t1 = time_best([&] { ffma_chain<<<blocks, threads>>>(out, 1.0000001f, 1e-9f, iters); }, 10);
  double t2 = time_best([&] { ffma2_chain<<<blocks, threads>>>(out, 1.0000001f, 1e-9f, iters); }, 10);
  double t3 = time_best([&] { ffma_tile<<<blocks, threads>>>(out, in, iters); }, 10);
  double t4 = time_best([&] { ffma2_tile<<<blocks, threads>>>(out, in, iters); }, 10);
  // sustained: run the best variant back-to-back for ~2 s
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  int n_sus = (int)(2000.0 / std::min(t1, t2)) + 1;
  CK(cudaEventRecord(e0));
  for (int i = 0; i < n_sus; ++i) {
    if (t2 < t1) ffma2_chain<<<blocks, threads>>>(out, 1.0000001f, 1e-9f, iters);
    else ffma_chain<<<blocks, threads>>>(out, 1.0000001f, 1e-9f, iters);
  }
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms_sus; CK(cudaEventElapsedTime(&ms_sus, e0, e1));
  CK(cudaGetLastError());
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d, "
         "\"ffma_chain_tflops\": %.2f, \"ffma2_chain_tflops\": %.2f, "
         "\"ffma_tile_tflops\": %.2f, \"ffma2_tile_tflops\": %.2f, "
         "\"sustained_tflops\": %.2f, \"nominal_tflops_at_max_clock\": %.2f}\n",
         p.name, sms, p.clockRate,
         fl_chain / t1 * 1e-9, fl_chain / t2 * 1e-9, fl_tile / t3 * 1e-9, fl_tile / t4 * 1e-9,
         fl_chain * n_sus / ms_sus * 1e-9, sms * 128.0 * 2.0 * p.clockRate * 1e-9);
  return 0;
}
