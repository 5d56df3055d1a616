// Reverse-engineers the shared-memory layout tcgen05.mma expects for an MN-major (transposed) tf32 A
// operand with SWIZZLE_NONE: one float of the A buffer is set to 1, B (K-major, known good) holds
// B[n][k] = k + 1, a single M=128,N=16,K=8 MMA runs, and the non-zero row m / value k+1 of D tell which
// (m, k) the hardware read from that byte offset.  Prints the map for a list of offsets and both
// descriptor field assignments.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../pinns_fluid_dynamics_b200/csrc/common.cuh"
#include "../pinns_fluid_dynamics_b200/csrc/umma.cuh"
using namespace pinn;
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

__global__ void __launch_bounds__(128) probe(uint32_t hot_off, uint32_t lbo, uint32_t sbo, int a_mn, float* __restrict__ C) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;            // 64 KB scanned region
  uint8_t* sB = smem + 65536;    // 16 x 8 K-major: (n,k) -> (n/8)*256 + (k/4)*128 + (n%8)*16 + (k%4)*4
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 65536 / 4; i += 128) reinterpret_cast<float*>(sA)[i] = 0.f;
  __syncthreads();
  if (tid == 0) *reinterpret_cast<float*>(sA + hot_off) = 1.0f;
  for (int i = tid; i < 16 * 8; i += 128) {
    const int n = i / 8, k = i % 8;
    *reinterpret_cast<float*>(sB + (n / 8) * 256 + (k / 4) * 128 + (n % 8) * 16 + (k % 4) * 4) = (float)(k + 1);
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) umma::tmem_alloc<32>(&tslot);
  umma::fence_proxy_async_smem();
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem = tslot;
  if (tid == 0) {
    const uint32_t idesc = umma::idesc_tf32(128, 16) | ((uint32_t)a_mn << 15);
    const uint64_t ad = umma::smem_desc((uint32_t)__cvta_generic_to_shared(sA), lbo, sbo);
    const uint64_t bd = umma::smem_desc((uint32_t)__cvta_generic_to_shared(sB), 128, 256);
    umma::mma_tf32_ss(tmem, ad, bd, idesc, 0);
    umma::commit(&bar);
  }
  mbar_wait(&bar, 0);
  umma::fence_after_thread_sync();
  float v[32];
  umma::tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16), v);
  C[warp * 32 + lane] = v[0];
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<32>(tmem);
}

int main() {
  float* dC; CK(cudaMalloc(&dC, 128 * 4));
  std::vector<float> C(128);
  const int smem = 65536 + 1024;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  struct D { uint32_t lbo, sbo; int a_mn; const char* name; };
  D descs[] = {{128, 1024, 0, "K-major  lbo=128  sbo=1024 (reference: expect m=(off/1024)*8+(off%128)/16, k=(off%1024)/128*4+(off%16)/4)"},
               {4096, 128, 1, "MN-major lbo=4096 sbo=128"}, {128, 4096, 1, "MN-major lbo=128  sbo=4096"},
               {256, 2048, 1, "MN-major lbo=256  sbo=2048"}, {2048, 256, 1, "MN-major lbo=2048 sbo=256"}};
  std::vector<uint32_t> offs;
  for (uint32_t o = 0; o < 64; o += 4) offs.push_back(o);
  for (uint32_t o = 64; o < 512; o += 16) offs.push_back(o);
  for (uint32_t o = 512; o <= 8192; o += 256) offs.push_back(o);
  offs.push_back(4096 + 16); offs.push_back(4096 + 128); offs.push_back(8192 + 4096);
  for (const D& d : descs) {
    printf("== %s\n", d.name);
    for (uint32_t o : offs) {
      probe<<<1, 128, smem>>>(o, d.lbo, d.sbo, d.a_mn, dC);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("off %u: CUDA error %s\n", o, cudaGetErrorString(e)); return 1; }
      CK(cudaMemcpy(C.data(), dC, 128 * 4, cudaMemcpyDeviceToHost));
      int hits = 0;
      for (int m = 0; m < 128; ++m)
        if (C[m] != 0.f) { printf("  off %5u -> m=%3d k=%d\n", o, m, (int)C[m] - 1); ++hits; }
      if (!hits) printf("  off %5u -> (not read)\n", o);
    }
  }
  return 0;
}
