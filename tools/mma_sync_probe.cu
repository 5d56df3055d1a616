// Throughput of the warp-level (legacy) tensor path on sm_100a: mma.sync.aligned.m16n8k8 tf32 -> f32, against the FFMA pipe.
// Would a 3xTF32 mma.sync GEMM beat the FP32 FFMA2 GEMMs of the fused H<=32 kernel?   usage: mma_sync_probe [warps_per_cta]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int NACC>
__global__ void mma_kernel(float* out, int iters) {
  float acc[NACC][4];
  unsigned a[4], b[2];
  for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1.0f + threadIdx.x * 1e-3f + i);
  for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(0.5f + threadIdx.x * 1e-3f + i);
#pragma unroll
  for (int n = 0; n < NACC; ++n)
    for (int i = 0; i < 4; ++i) acc[n][i] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int n = 0; n < NACC; ++n) mma_tf32(acc[n], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int n = 0; n < NACC; ++n)
    for (int i = 0; i < 4; ++i) s += acc[n][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main(int argc, char** argv) {
  int dev = 0;
  cudaSetDevice(dev);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, dev);
  const int sms = prop.multiProcessorCount;
  float* out;
  cudaMalloc(&out, sizeof(float) * sms * 1024 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    for (int nacc : {4, 8, 16}) {
      auto launch = [&](int it) {
        if (nacc == 4) mma_kernel<4><<<sms, warps * 32>>>(out, it);
        else if (nacc == 8) mma_kernel<8><<<sms, warps * 32>>>(out, it);
        else mma_kernel<16><<<sms, warps * 32>>>(out, it);
      };
      launch(100);
      cudaDeviceSynchronize();
      cudaEventRecord(e0);
      launch(iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double flops = 2.0 * 16 * 8 * 8 * (double)nacc * iters * warps * sms;
      printf("warps/SM %2d  independent accumulators %2d : %.1f TFLOP/s (m16n8k8 tf32), %.2f mma/clk/SM at %d MHz nominal\n", warps, nacc,
             flops / ms * 1e-9, (double)nacc * iters * warps / (ms * 1e-3 * prop.clockRate * 1e3), prop.clockRate / 1000);
    }
  }
  printf("cudaGetLastError: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
