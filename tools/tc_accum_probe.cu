// How accurate is the tensor core's FP32 accumulation for the 3-pass TF32 split?  One CTA computes
// C[128x128] = A[128xK] * Bt[128xK]^T (K = 128) with hi/lo operands that are exact TF32 values (rna split),
// in several accumulation arrangements, and reports the error against float64:
//   mode 0  single accumulator, per k-step (lo*hi, hi*lo, hi*hi)                    [engine v1]
//   mode 1  hi*hi only                         vs  float64 sum of hi*hi products   (pure accumulation error)
//   mode 2  big accumulator (hi*hi) + small accumulator (lo*hi, hi*lo), added in FP32 by the epilogue
//   mode 3  as mode 2 plus lo*lo in the small accumulator
// for random-sign data and for all-positive data (where a one-sided truncation shows up as a bias).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include "../pinns_fluid_dynamics_b200/csrc/common.cuh"
#include "../pinns_fluid_dynamics_b200/csrc/umma.cuh"
using namespace pinn;
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)
constexpr int M = 128, N = 128, K = 128;

__global__ void __launch_bounds__(128) probe(const float* __restrict__ A, const float* __restrict__ Bt, float* __restrict__ C, int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // K = 128 does not fit four 64 KB operand tiles: process K in two halves of 64 (4 x 32 KB)
  uint8_t* sAh = smem; uint8_t* sAl = sAh + 32768; uint8_t* sBh = sAl + 32768; uint8_t* sBl = sBh + 32768;
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int KH = 64;
  constexpr uint32_t SBO = umma::sbo_for_k(KH);
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) umma::tmem_alloc<256>(&tslot);
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem = tslot;
  uint32_t acc_big = 0, acc_small = 0;
  for (int half = 0; half < 2; ++half) {
    for (int idx = tid; idx < M * KH; idx += 128) {
      const int r = idx / KH, k = idx % KH;
      float ah, al, bh, bl;
      umma::split_tf32_rn(A[(size_t)r * K + half * KH + k], ah, al);
      umma::split_tf32_rn(Bt[(size_t)r * K + half * KH + k], bh, bl);
      const uint32_t off = umma::tile_offset(r, k, SBO);
      *reinterpret_cast<float*>(sAh + off) = ah; *reinterpret_cast<float*>(sAl + off) = al;
      *reinterpret_cast<float*>(sBh + off) = bh; *reinterpret_cast<float*>(sBl + off) = bl;
    }
    umma::fence_proxy_async_smem();
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    if (tid == 0) {
      const uint32_t idesc = umma::idesc_tf32(M, N);
      const uint32_t aH = (uint32_t)__cvta_generic_to_shared(sAh), aL = (uint32_t)__cvta_generic_to_shared(sAl);
      const uint32_t bH = (uint32_t)__cvta_generic_to_shared(sBh), bL = (uint32_t)__cvta_generic_to_shared(sBl);
      const uint32_t dbig = tmem, dsmall = (mode >= 2) ? tmem + 128 : tmem;
      for (int ks = 0; ks < KH / 8; ++ks) {
        const uint64_t ah = umma::smem_desc(aH + ks * 256, 128, SBO), al = umma::smem_desc(aL + ks * 256, 128, SBO);
        const uint64_t bh = umma::smem_desc(bH + ks * 256, 128, SBO), bl = umma::smem_desc(bL + ks * 256, 128, SBO);
        if (mode != 1) {
          uint32_t& as = (mode >= 2) ? acc_small : acc_big;
          umma::mma_tf32_ss(dsmall, al, bh, idesc, as); as = 1;
          umma::mma_tf32_ss(dsmall, ah, bl, idesc, 1);
          if (mode == 3) umma::mma_tf32_ss(dsmall, al, bl, idesc, 1);
        }
        umma::mma_tf32_ss(dbig, ah, bh, idesc, acc_big); acc_big = 1;
      }
      umma::commit(&bar);
    }
    mbar_wait(&bar, half);
    umma::fence_after_thread_sync();
    __syncthreads();
  }
#pragma unroll 1
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32], w[32];
    umma::tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    if (mode >= 2) {
      umma::tmem_ld_32x32(tmem + 128 + ((uint32_t)(warp * 32) << 16) + c0, w);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += w[j];
    }
    const int row = warp * 32 + lane;
#pragma unroll
    for (int j = 0; j < 32; ++j) C[(size_t)row * N + c0 + j] = v[j];
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<256>(tmem);
}

static float rna(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

int main() {
  std::vector<float> A((size_t)M * K), Bt((size_t)N * K), C((size_t)M * N);
  float *dA, *dB, *dC;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, Bt.size() * 4)); CK(cudaMalloc(&dC, C.size() * 4));
  const int smem = 4 * 32768;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int positive = 0; positive < 2; ++positive) {
    srand(7);
    for (auto& v : A) v = positive ? (float)rand() / RAND_MAX : (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto& v : Bt) v = positive ? (float)rand() / RAND_MAX : (float)rand() / RAND_MAX * 2.f - 1.f;
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bt.data(), Bt.size() * 4, cudaMemcpyHostToDevice));
    // FP32 FMA reference error for comparison
    double fe = 0, fn = 0, fbias = 0;
    for (int i = 0; i < M; ++i)
      for (int j = 0; j < N; ++j) {
        double ex = 0; float f = 0.f;
        for (int k = 0; k < K; ++k) { ex += (double)A[(size_t)i * K + k] * Bt[(size_t)j * K + k]; f = fmaf(A[(size_t)i * K + k], Bt[(size_t)j * K + k], f); }
        fe += (f - ex) * (f - ex); fn += ex * ex; fbias += (f - ex) / fabs(ex);
      }
    printf("[%s data] FP32 FMA chain: rel L2 err %.3e, mean signed rel err %.3e\n", positive ? "positive" : "random-sign", sqrt(fe / fn), fbias / (M * N));
    for (int mode = 0; mode < 4; ++mode) {
      probe<<<1, 128, smem>>>(dA, dB, dC, mode);
      CK(cudaGetLastError());
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
      double e2 = 0, n2 = 0, bias = 0;
      for (int i = 0; i < M; ++i)
        for (int j = 0; j < N; ++j) {
          double ex = 0;
          for (int k = 0; k < K; ++k) {
            const float a = A[(size_t)i * K + k], b = Bt[(size_t)j * K + k];
            if (mode == 1) ex += (double)rna(a) * (double)rna(b);
            else ex += (double)a * b;
          }
          const double c = C[(size_t)i * N + j];
          e2 += (c - ex) * (c - ex); n2 += ex * ex; bias += (c - ex) / fabs(ex);
        }
      printf("[%s data] mode %d: rel L2 err %.3e, mean signed rel err %.3e\n", positive ? "positive" : "random-sign", mode, sqrt(e2 / n2), bias / (M * N));
    }
  }
  return 0;
}
