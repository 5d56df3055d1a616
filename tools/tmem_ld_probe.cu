// tcgen05.ld throughput: W warps (quarter = warp % 4) each issue `iters` 32x32b.x32 loads (4 KB per warp
// instruction) from their own TMEM lane quarter; optionally one extra warp keeps the tensor pipe busy with
// N=256 tf32 MMAs meanwhile.  Reports bytes per SM cycle.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include "../pinns_fluid_dynamics_b200/csrc/common.cuh"
#include "../pinns_fluid_dynamics_b200/csrc/umma.cuh"
using namespace pinn;
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

__global__ void __launch_bounds__(544) k(int ld_warps, int iters, int with_mma, long long* cyc, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  __shared__ volatile int done;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (160 * 1024) / 16; i += blockDim.x) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid == 0) { mbar_init(&bar, 1); done = 0; fence_barrier_init(); }
  if (warp == 0) umma::tmem_alloc<512>(&tslot);
  umma::fence_proxy_async_smem();
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem = tslot;
  if (warp < ld_warps) {
    float acc = 0.f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      float v[32];
      umma::tmem_ld_32x32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + ((it * 32) & 255), v);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc += v[j];
    }
    const long long t1 = clock64();
    if (lane == 0) cyc[blockIdx.x * 16 + warp] = t1 - t0;
    if (acc == 123.456f) sink[0] = acc;
    if (warp == 0 && lane == 0) done = 1;
  } else if (warp == 16 && lane == 0 && with_mma) {
    const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t idesc = umma::idesc_tf32(128, 256);
    int n = 0;
    while (!done && n < 200000) {
      for (int u = 0; u < 8; ++u) {
        const uint64_t ad = umma::smem_desc(s0 + u * 256, 128, 4096), bd = umma::smem_desc(s0 + 65536 + u * 256, 128, 256);
        umma::mma_tf32_ss(tmem + 256, ad, bd, idesc, 1);
      }
      umma::commit(&bar);
      mbar_wait(&bar, n & 1);
      ++n;
    }
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<512>(tmem);
}

int main() {
  long long* dc; float* sink; CK(cudaMalloc(&dc, 148 * 16 * 8)); CK(cudaMalloc(&sink, 4));
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  long long hc[148 * 16];
  const int iters = 4096;
  for (int with_mma = 0; with_mma < 2; ++with_mma)
    for (int w : {1, 4, 8, 16}) {
      CK(cudaMemset(dc, 0, 148 * 16 * 8));
      k<<<148, 544, 160 * 1024>>>(w, iters, with_mma, dc, sink);
      CK(cudaGetLastError());
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(hc, dc, sizeof(hc), cudaMemcpyDeviceToHost));
      double mx = 0; for (int i = 0; i < w; ++i) mx = hc[i] > mx ? (double)hc[i] : mx;
      printf("ld warps %2d, mma %d: %.1f cycles per x32 load per warp, %.1f B/cycle/SM\n", w, with_mma, mx / iters, (double)w * iters * 4096 / mx);
    }
  return 0;
}
