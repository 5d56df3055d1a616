// tcgen05 issue-rate probe with a TIGHT issue loop: the whole MMA warp runs the loop, one lane chosen by elect.sync issues
// (the compiler then keeps descriptors in uniform registers and emits bare UTCHMMA back to back; an `if (tid == X)` issuer
// wraps every UTCHMMA in an ELECT / BRA.U.ANY loop that costs ~50 cycles per MMA -- tools/tc_probe2.cu measured that).
// Shapes: the forward / input-adjoint GEMM of the H = 32 fused kernel (A in tensor memory, weights K-major in shared
// memory) and the weight-gradient form (bf16, both operands MN-major).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../pinns_fluid_dynamics_b200/csrc/common.cuh"
#include "../pinns_fluid_dynamics_b200/csrc/umma.cuh"
using namespace pinn;
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mma_tf32_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a),
               "l"(b), "r"(idesc)
               : "memory");
}
__device__ __forceinline__ void mma_tf32_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a),
               "l"(b), "r"(idesc)
               : "memory");
}
__device__ __forceinline__ void mma_f16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a),
               "l"(b), "r"(idesc)
               : "memory");
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a),
               "l"(b), "r"(idesc)
               : "memory");
}

// KIND 0: TS tf32 M128 N32, channel-outer (12 MMAs per D tile), 60 / commit
//      1: TS tf32 M128 N32, channel-inner (D tile changes every MMA)
//      2: TS tf32 M128 N64 then N32 per k-step (hi x [W_hi;W_lo], lo x W_hi), D tiles of 64 columns, 2 D slots
//      3: SS bf16 MN-major M64 N64 K16 (stacked weight-gradient form), 40 / commit, one accumulator
//      4: SS bf16 MN-major M64 N32 K16, 120 / commit
//      5: TS tf32 M128 N16 | 6: TS tf32 M128 N128 | 7: TS tf32 M128 N256 | 8: SS tf32 M128 N32 (A from shared memory)
//      9: TS bf16 K-major M128 N32 K16, channel-outer, 6 passes x 2 k-steps
//     10: TS tf32 M128 N32 channel-outer, D tile changes every 4 MMAs (k-step-major within a pass)
//     11: TS tf32 M64 N32
template <int KIND>
__global__ void __launch_bounds__(64) rate3(int stages, long long* cycles, long long* n_mma) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (96 * 1024) / 16; i += 64) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
  if (warp == 0) umma::tmem_alloc<512>(&tslot);
  umma::fence_proxy_async_smem();
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem = tslot;
  if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(smem);
    const uint64_t bh0 = umma::smem_desc(s0, 128, 1024), bl0 = umma::smem_desc(s0 + 4096, 128, 1024);
    const uint64_t ah0 = umma::smem_desc(s0 + 16384, 128, 1024), al0 = umma::smem_desc(s0 + 32768, 128, 1024);
    const uint64_t bst0 = umma::smem_desc(s0, 128, 1024);                 // stacked [W_hi; W_lo]: 64 rows, SBO 1024
    const uint64_t gA64 = umma::smem_desc(s0 + 49152, 1024, 128), gZ64 = umma::smem_desc(s0 + 49152 + 22528, 1024, 128);
    const uint64_t gA32 = umma::smem_desc(s0 + 49152, 512, 128), gZ32 = umma::smem_desc(s0 + 49152 + 22528, 512, 128);
    const uint64_t bb0 = umma::smem_desc(s0, 128, 512);
    long long count = 0;
    const long long t0 = clock64();
    for (int s = 0; s < stages; ++s) {
      if (s >= 2) mbar_wait(&bar[s & 1], ((s >> 1) - 1) & 1);
      if (leader) {
        if constexpr (KIND == 0 || KIND == 5 || KIND == 6 || KIND == 7 || KIND == 8 || KIND == 11) {
          constexpr int N = KIND == 5 ? 16 : KIND == 6 ? 128 : KIND == 7 ? 256 : 32;
          constexpr int TILES = N >= 128 ? 1 : 5;
          constexpr uint32_t idesc = umma::idesc_tf32(KIND == 11 ? 64 : 128, N);
#pragma unroll
          for (int c = 0; c < TILES; ++c)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t d = tmem + (uint32_t)(c * (N > 32 ? 0 : 32));
              const uint32_t ah = tmem + 256 + (uint32_t)((c % 4) * 32 + ks * 8), al = ah + 128;
              if constexpr (KIND == 8) {
                mma_tf32_ss(d, al0 + (uint64_t)(ks * 16), bh0 + (uint64_t)(ks * 16), idesc);
                mma_tf32_ss(d, ah0 + (uint64_t)(ks * 16), bl0 + (uint64_t)(ks * 16), idesc);
                mma_tf32_ss(d, ah0 + (uint64_t)(ks * 16), bh0 + (uint64_t)(ks * 16), idesc);
              } else {
                mma_tf32_ts(d, al, bh0 + (uint64_t)(ks * 16), idesc);
                mma_tf32_ts(d, ah, bl0 + (uint64_t)(ks * 16), idesc);
                mma_tf32_ts(d, ah, bh0 + (uint64_t)(ks * 16), idesc);
              }
            }
          count += TILES * 12;
        } else if constexpr (KIND == 1) {
          constexpr uint32_t idesc = umma::idesc_tf32(128, 32);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
#pragma unroll
            for (int p = 0; p < 3; ++p)
#pragma unroll
              for (int c = 0; c < 5; ++c) {
                const uint32_t ah = tmem + 256 + (uint32_t)((c % 4) * 32 + ks * 8), al = ah + 128;
                mma_tf32_ts(tmem + (uint32_t)(c * 32), p == 0 ? al : ah, (p == 1 ? bl0 : bh0) + (uint64_t)(ks * 16), idesc);
              }
          count += 60;
        } else if constexpr (KIND == 10) {
          constexpr uint32_t idesc = umma::idesc_tf32(128, 32);
#pragma unroll
          for (int c = 0; c < 5; ++c)
#pragma unroll
            for (int p = 0; p < 3; ++p)
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint32_t ah = tmem + 256 + (uint32_t)((c % 4) * 32 + ks * 8), al = ah + 128;
                mma_tf32_ts(tmem + (uint32_t)(c * 32), p == 0 ? al : ah, (p == 1 ? bl0 : bh0) + (uint64_t)(ks * 16), idesc);
              }
          count += 60;
        } else if constexpr (KIND == 2) {
          constexpr uint32_t i64 = umma::idesc_tf32(128, 64), i32 = umma::idesc_tf32(128, 32);
#pragma unroll
          for (int c = 0; c < 5; ++c)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t d = tmem + (uint32_t)((c & 1) * 64);
              const uint32_t ah = tmem + 256 + (uint32_t)((c % 4) * 32 + ks * 8), al = ah + 128;
              mma_tf32_ts(d, al, bh0 + (uint64_t)(ks * 16), i32);
              mma_tf32_ts(d, ah, bst0 + (uint64_t)(ks * 16), i64);
            }
          count += 40;
        } else if constexpr (KIND == 3) {
          constexpr uint32_t idesc = idesc_bf16(64, 64, 1, 1);
#pragma unroll 8
          for (int ks = 0; ks < 40; ++ks)
            mma_f16_ss(tmem + 448, gA64 + (uint64_t)((ks % 10) * 128), gZ64 + (uint64_t)((ks % 10) * 128), idesc);
          count += 40;
        } else if constexpr (KIND == 4) {
          constexpr uint32_t idesc = idesc_bf16(64, 32, 1, 1);
#pragma unroll 8
          for (int ks = 0; ks < 120; ++ks)
            mma_f16_ss(tmem + 448, gA32 + (uint64_t)((ks % 20) * 64), gZ32 + (uint64_t)((ks % 20) * 64), idesc);
          count += 120;
        } else if constexpr (KIND == 9) {
          constexpr uint32_t idesc = idesc_bf16(128, 32, 0, 0);
#pragma unroll
          for (int c = 0; c < 5; ++c)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
#pragma unroll
              for (int p = 0; p < 6; ++p)
                mma_f16_ts(tmem + (uint32_t)(c * 32), tmem + 256 + (uint32_t)(c * 48 + (p % 3) * 16 + ks * 8), bb0 + (uint64_t)(ks * 16 + (p & 1) * 128), idesc);
          count += 60;
        }
        umma::commit(&bar[s & 1]);
      }
      __syncwarp();
    }
    const int last = stages - 1;
    mbar_wait(&bar[last & 1], (last >> 1) & 1);
    if (stages >= 2) mbar_wait(&bar[(last - 1) & 1], ((last - 1) >> 1) & 1);
    if (leader) {
      cycles[blockIdx.x] = clock64() - t0;
      n_mma[blockIdx.x] = count;
    }
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<512>(tmem);
}

template <int KIND>
static void run(const char* name, int sms, long long* dc, long long* dn) {
  const int rsmem = 96 * 1024;
  CK(cudaFuncSetAttribute(rate3<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, rsmem));
  const int stages = 1000;
  rate3<KIND><<<sms, 64, rsmem>>>(16, dc, dn);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  rate3<KIND><<<sms, 64, rsmem>>>(stages, dc, dn);
  CK(cudaEventRecord(e1));
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kind %d: CUDA error %s\n", KIND, cudaGetErrorString(e)); exit(1); }
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> hc(sms), hn(sms);
  CK(cudaMemcpy(hc.data(), dc, 8 * sms, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hn.data(), dn, 8 * sms, cudaMemcpyDeviceToHost));
  double cyc = 0; for (int i = 0; i < sms; ++i) cyc += (double)hc[i]; cyc /= sms;
  printf("%-64s: %.3f ms, %.1f SM cycles/MMA, %.0f cycles per stage of %lld, clock %.0f MHz\n", name, ms, cyc / (double)hn[0], cyc / stages,
         hn[0] / stages, cyc / (ms * 1e-3) * 1e-6);
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  long long *dc, *dn;
  CK(cudaMalloc(&dc, 8 * 256)); CK(cudaMalloc(&dn, 8 * 256));
  run<0>("TS tf32 M128 N32 K8, 12 MMAs per D tile, 5 tiles", sms, dc, dn);
  run<10>("TS tf32 M128 N32 K8, pass-major within a D tile", sms, dc, dn);
  run<1>("TS tf32 M128 N32 K8, D tile changes every MMA", sms, dc, dn);
  run<2>("TS tf32 M128: lo x W_hi (N32) + hi x [W_hi;W_lo] (N64) per k-step", sms, dc, dn);
  run<8>("SS tf32 M128 N32 K8 (A from shared memory)", sms, dc, dn);
  run<11>("TS tf32 M64 N32 K8", sms, dc, dn);
  run<5>("TS tf32 M128 N16", sms, dc, dn);
  run<6>("TS tf32 M128 N128", sms, dc, dn);
  run<7>("TS tf32 M128 N256", sms, dc, dn);
  run<9>("TS bf16 K-major M128 N32 K16, 6 passes x 2 k-steps per D tile", sms, dc, dn);
  run<3>("SS bf16 MN-major M64 N64 K16 (stacked weight gradient)", sms, dc, dn);
  run<4>("SS bf16 MN-major M64 N32 K16", sms, dc, dn);
  return 0;
}
