// Unit test of the tcgen05 kind::tf32 building blocks (descriptors, TMEM, 3-pass split):
// C[M x N] = A[M x K] * Bt[N x K]^T, M = 128 per CTA, N = 64, K = 128.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../pinns_fluid_dynamics_b200/csrc/common.cuh"
#include "../pinns_fluid_dynamics_b200/csrc/umma.cuh"
using namespace pinn;
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

constexpr int M = 128, N = 64, K = 128;

// mode 0: single pass on raw inputs; 1: single pass on explicit hi; 2: three-pass split
__global__ void __launch_bounds__(128) gemm_test(const float* __restrict__ A, const float* __restrict__ Bt, float* __restrict__ C, int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr uint32_t SBO = umma::sbo_for_k(K);
  uint8_t* sAh = smem;                       // 128 x 128 x 4 = 64 KB
  uint8_t* sAl = sAh + M * K * 4;
  uint8_t* sBh = sAl + M * K * 4;            // 64 x 128 x 4 = 32 KB
  uint8_t* sBl = sBh + N * K * 4;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sBl + N * K * 4);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* Ag = A + (size_t)blockIdx.x * M * K;
  // producer: 8 rows x 4 k-chunks per warp instruction
  for (int it = 0; it < (M * K / 4) / 128; ++it) {
    const int g = it * 4 + warp;                 // group of (8 rows x 4 chunks): 16 row-groups x 8 chunk-groups
    const int rg = g >> 3, cg = g & 7;
    const int r = rg * 8 + (lane & 7), kc = cg * 4 + (lane >> 3);
    const float4 v = *reinterpret_cast<const float4*>(Ag + (size_t)r * K + kc * 4);
    float4 h, l;
    if (mode == 0) { h = v; l = make_float4(0, 0, 0, 0); }
    else if (mode == 3) { umma::split_tf32_rn(v.x, h.x, l.x); umma::split_tf32_rn(v.y, h.y, l.y); umma::split_tf32_rn(v.z, h.z, l.z); umma::split_tf32_rn(v.w, h.w, l.w); }
    else { umma::split_tf32(v.x, h.x, l.x); umma::split_tf32(v.y, h.y, l.y); umma::split_tf32(v.z, h.z, l.z); umma::split_tf32(v.w, h.w, l.w); }
    const uint32_t off = umma::tile_offset(r, kc * 4, SBO);
    *reinterpret_cast<float4*>(sAh + off) = h;
    *reinterpret_cast<float4*>(sAl + off) = l;
  }
  for (int it = 0; it < (N * K / 4) / 128; ++it) {
    const int g = it * 4 + warp;
    const int rg = g >> 3, cg = g & 7;
    const int r = rg * 8 + (lane & 7), kc = cg * 4 + (lane >> 3);
    const float4 v = *reinterpret_cast<const float4*>(Bt + (size_t)r * K + kc * 4);
    float4 h, l;
    if (mode == 0) { h = v; l = make_float4(0, 0, 0, 0); }
    else if (mode == 3) { umma::split_tf32_rn(v.x, h.x, l.x); umma::split_tf32_rn(v.y, h.y, l.y); umma::split_tf32_rn(v.z, h.z, l.z); umma::split_tf32_rn(v.w, h.w, l.w); }
    else { umma::split_tf32(v.x, h.x, l.x); umma::split_tf32(v.y, h.y, l.y); umma::split_tf32(v.z, h.z, l.z); umma::split_tf32(v.w, h.w, l.w); }
    const uint32_t off = umma::tile_offset(r, kc * 4, SBO);
    *reinterpret_cast<float4*>(sBh + off) = h;
    *reinterpret_cast<float4*>(sBl + off) = l;
  }
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) umma::tmem_alloc<64>(tslot);
  umma::fence_proxy_async_smem();
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem = *tslot;
  if (tid == 0) {
    const uint32_t idesc = umma::idesc_tf32(M, N);
    const uint32_t aH = (uint32_t)__cvta_generic_to_shared(sAh), aL = (uint32_t)__cvta_generic_to_shared(sAl);
    const uint32_t bH = (uint32_t)__cvta_generic_to_shared(sBh), bL = (uint32_t)__cvta_generic_to_shared(sBl);
    const int n_pass = mode >= 2 ? 3 : 1;
    uint32_t acc = 0;
    for (int p = 0; p < n_pass; ++p) {
      const uint32_t a0 = p == 0 ? aL : aH, b0 = p == 1 ? bL : bH;   // lo*hi, hi*lo, hi*hi (small terms first)
      for (int ks = 0; ks < K / 8; ++ks) {
        const uint64_t ad = umma::smem_desc(a0 + ks * 2 * umma::kLBO, umma::kLBO, SBO);
        const uint64_t bd = umma::smem_desc(b0 + ks * 2 * umma::kLBO, umma::kLBO, SBO);
        umma::mma_tf32_ss(tmem, ad, bd, idesc, acc);
        acc = 1;
      }
    }
    umma::commit(bar);
  }
  mbar_wait(bar, 0);
  umma::fence_after_thread_sync();
  float* Cg = C + (size_t)blockIdx.x * M * N;
#pragma unroll 1
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    umma::tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    const int row = warp * 32 + lane;
#pragma unroll
    for (int j = 0; j < 32; j += 4)
      *reinterpret_cast<float4*>(Cg + (size_t)row * N + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<64>(tmem);
}

static float trunc_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float rna_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

int main() {
  const int tiles = 4;
  std::vector<float> A((size_t)tiles * M * K), Bt((size_t)N * K), C((size_t)tiles * M * N);
  srand(1);
  for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& v : Bt) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  float *dA, *dB, *dC;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, Bt.size() * 4)); CK(cudaMalloc(&dC, C.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bt.data(), Bt.size() * 4, cudaMemcpyHostToDevice));
  const int smem = 2 * M * K * 4 + 2 * N * K * 4 + 64;
  CK(cudaFuncSetAttribute(gemm_test, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int mode = 0; mode < 4; ++mode) {
    CK(cudaMemset(dC, 0, C.size() * 4));
    gemm_test<<<tiles, 128, smem>>>(dA, dB, dC, mode);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
    double e_exact = 0, e_trunc = 0, e_rna = 0, ref_norm = 0;
    for (int t = 0; t < tiles; ++t)
      for (int i = 0; i < M; ++i)
        for (int j = 0; j < N; ++j) {
          double ex = 0, tr = 0, rn = 0;
          for (int k = 0; k < K; ++k) {
            const float a = A[((size_t)t * M + i) * K + k], b = Bt[(size_t)j * K + k];
            ex += (double)a * b; tr += (double)trunc_tf32(a) * trunc_tf32(b); rn += (double)rna_tf32(a) * rna_tf32(b);
          }
          const double c = C[((size_t)t * M + i) * N + j];
          e_exact += (c - ex) * (c - ex); e_trunc += (c - tr) * (c - tr); e_rna += (c - rn) * (c - rn); ref_norm += ex * ex;
        }
    printf("mode %d (%s): rel L2 err vs exact %.3e, vs truncated-input product %.3e, vs rna-input product %.3e\n", mode,
           mode == 0 ? "1 pass raw" : mode == 1 ? "1 pass explicit hi" : mode == 2 ? "3 pass trunc split" : "3 pass rna split", sqrt(e_exact / ref_norm),
           sqrt(e_trunc / ref_norm), sqrt(e_rna / ref_norm));
  }
  printf("C[0][0..3] = %f %f %f %f\n", C[0], C[1], C[2], C[3]);
  return 0;
}
