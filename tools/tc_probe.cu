// tcgen05 kind::tf32 probes for the layered tensor-core engine (run on a B200 through gpurun):
//   part 1  correctness of the operand forms the engine relies on
//           (a) single pass on RAW fp32 operands (does the tensor core truncate the low 13 bits itself?)
//           (b) 3-pass split, K-major A and B (forward / backward layer GEMMs)
//           (c) MN-major A and B, no swizzle, both (LBO,SBO) assignments (weight-gradient GEMM)
//           (d) A operand from tensor memory (tcgen05.st -> tcgen05.mma [d],[a],b-desc)
//   part 2  throughput of the issue loop the kernels use: 148 CTAs x 1 issuing thread, N = 128,
//           SS K-major / SS MN-major / TS, with and without 4 warps writing shared memory concurrently.
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

// instruction descriptor with selectable majors (bit 15: A MN-major, bit 16: B MN-major)
__host__ __device__ constexpr uint32_t idesc_tf32_major(int m, int n, int a_mn, int b_mn) {
  return umma::idesc_tf32(m, n) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16);
}

__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint64_t with_swizzle128(uint64_t d) { return d | ((uint64_t)2 << 61); }

__device__ __forceinline__ void tmem_st_32x32_x8(uint32_t taddr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
               "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}

// modes: 0 raw single pass (K-major) | 1 3-pass K-major | 2 3-pass MN-major cand.1 (lbo = k-block stride, sbo = mn-block stride)
//        3 3-pass MN-major cand.2 (swapped) | 4 3-pass, A from TMEM (hi and lo regions), B K-major
// K-major smem layout  : (r,k) -> (r/8)*SBO + (k/4)*128 + (r%8)*16 + (k%4)*4,  SBO = (K/4)*128
// MN-major smem layout : (r,k) -> (k/8)*KB  + (r/4)*128 + (k%8)*16 + (r%4)*4,  KB  = (R/4)*128  (R rows = M or N)
__global__ void __launch_bounds__(128) gemm_probe(const float* __restrict__ A, const float* __restrict__ Bt, float* __restrict__ C, int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sAh = smem;                       // 64 KB each
  uint8_t* sAl = sAh + M * K * 4;
  uint8_t* sBh = sAl + M * K * 4;
  uint8_t* sBl = sBh + N * K * 4;            // total 256 KB is too much: lo buffers only hold data in 3-pass modes
  (void)sBl;
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool mn = (mode == 2 || mode == 3);
  constexpr uint32_t SBO_K = umma::sbo_for_k(K);
  // fill: element (r,k) of A / Bt -> smem
  for (int idx = tid; idx < M * K; idx += 128) {
    const int r = idx / K, k = idx % K;
    const float a = A[idx], b = Bt[idx];
    float ah, al, bh, bl;
    if (mode == 0) { ah = a; al = 0.f; bh = b; bl = 0.f; }
    else { umma::split_tf32(a, ah, al); umma::split_tf32(b, bh, bl); }
    uint32_t off;
    if (!mn) off = umma::tile_offset(r, k, SBO_K);
    else off = (uint32_t)(k >> 3) * (uint32_t)(M / 4) * 128u + (uint32_t)(r >> 2) * 128u + (uint32_t)(k & 7) * 16u + (uint32_t)(r & 3) * 4u;
    *reinterpret_cast<float*>(sAh + off) = ah;
    *reinterpret_cast<float*>(sAl + off) = al;
    *reinterpret_cast<float*>(sBh + off) = bh;
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) umma::tmem_alloc<512>(&tslot);
  umma::fence_proxy_async_smem();
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem = tslot;
  const uint32_t tA_hi = tmem + 128, tA_lo = tmem + 256;
  if (mode == 4) {   // thread (warp, lane) owns row warp*32+lane: write its 128 K-values into TMEM columns
    const int row = warp * 32 + lane;
    for (int k0 = 0; k0 < K; k0 += 8) {
      float h[8], l[8];
      for (int j = 0; j < 8; ++j) umma::split_tf32(A[(size_t)row * K + k0 + j], h[j], l[j]);
      tmem_st_32x32_x8(tA_hi + ((uint32_t)(warp * 32) << 16) + k0, h);
      tmem_st_32x32_x8(tA_lo + ((uint32_t)(warp * 32) << 16) + k0, l);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
  }
  // B lo lives in the A-lo buffer's twin only when needed: reuse sBl region? (smem is 3 x 64 KB) -> B lo is
  // recomputed into sAl for the pass that needs it in K-major mode 1 after A-lo has been consumed.
  if (tid == 0) {
    const uint32_t aH = (uint32_t)__cvta_generic_to_shared(sAh), aL = (uint32_t)__cvta_generic_to_shared(sAl);
    const uint32_t bH = (uint32_t)__cvta_generic_to_shared(sBh);
    uint32_t idesc = umma::idesc_tf32(M, N);
    if (mn) idesc = idesc_tf32_major(M, N, 1, 1);
    // pass list: (A_lo,B_hi), (A_hi,B_hi)   [B_lo pass issued in a second phase below]
    uint32_t acc = 0;
    const int n_pass = (mode == 0) ? 1 : 2;
    for (int p = 0; p < n_pass; ++p) {
      const bool a_is_lo = (n_pass == 2 && p == 0);
      for (int ks = 0; ks < K / 8; ++ks) {
        uint64_t ad, bd;
        if (!mn) {
          ad = umma::smem_desc((a_is_lo ? aL : aH) + ks * 2 * umma::kLBO, umma::kLBO, SBO_K);
          bd = umma::smem_desc(bH + ks * 2 * umma::kLBO, umma::kLBO, SBO_K);
        } else {
          const uint32_t kb = (uint32_t)(M / 4) * 128u;   // stride between 8-wide k blocks
          const uint32_t lbo = (mode == 2) ? kb : 128u, sbo = (mode == 2) ? 128u : kb;
          ad = umma::smem_desc((a_is_lo ? aL : aH) + ks * kb, lbo, sbo);
          bd = umma::smem_desc(bH + ks * kb, lbo, sbo);
        }
        if (mode == 4) mma_tf32_ts(tmem, (a_is_lo ? tA_lo : tA_hi) + ks * 8, bd, idesc, acc);
        else umma::mma_tf32_ss(tmem, ad, bd, idesc, acc);
        acc = 1;
      }
    }
    umma::commit(&bar);
  }
  mbar_wait(&bar, 0);
  umma::fence_after_thread_sync();
  if (mode != 0) {
    // third pass A_hi x B_lo: overwrite the A-lo buffer with B-lo (A-lo has been consumed)
    __syncthreads();
    for (int idx = tid; idx < M * K; idx += 128) {
      const int r = idx / K, k = idx % K;
      float bh, bl;
      umma::split_tf32(Bt[idx], bh, bl);
      uint32_t off;
      if (!mn) off = umma::tile_offset(r, k, SBO_K);
      else off = (uint32_t)(k >> 3) * (uint32_t)(M / 4) * 128u + (uint32_t)(r >> 2) * 128u + (uint32_t)(k & 7) * 16u + (uint32_t)(r & 3) * 4u;
      *reinterpret_cast<float*>(sAl + off) = bl;
    }
    umma::fence_proxy_async_smem();
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    if (tid == 0) {
      const uint32_t aH = (uint32_t)__cvta_generic_to_shared(sAh), bL = (uint32_t)__cvta_generic_to_shared(sAl);
      uint32_t idesc = umma::idesc_tf32(M, N);
      if (mn) idesc = idesc_tf32_major(M, N, 1, 1);
      for (int ks = 0; ks < K / 8; ++ks) {
        uint64_t ad, bd;
        if (!mn) {
          ad = umma::smem_desc(aH + ks * 2 * umma::kLBO, umma::kLBO, SBO_K);
          bd = umma::smem_desc(bL + ks * 2 * umma::kLBO, umma::kLBO, SBO_K);
        } else {
          const uint32_t kb = (uint32_t)(M / 4) * 128u;
          const uint32_t lbo = (mode == 2) ? kb : 128u, sbo = (mode == 2) ? 128u : kb;
          ad = umma::smem_desc(aH + ks * kb, lbo, sbo);
          bd = umma::smem_desc(bL + ks * kb, lbo, sbo);
        }
        if (mode == 4) mma_tf32_ts(tmem, tA_hi + ks * 8, bd, idesc, 1);
        else umma::mma_tf32_ss(tmem, ad, bd, idesc, 1);
      }
      umma::commit(&bar);
    }
    mbar_wait(&bar, 1);
    umma::fence_after_thread_sync();
  }
#pragma unroll 1
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    umma::tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    const int row = warp * 32 + lane;
#pragma unroll
    for (int j = 0; j < 32; j += 4)
      *reinterpret_cast<float4*>(C + (size_t)row * N + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<512>(tmem);
}

// ---- throughput probe ------------------------------------------------------------------------------
// kind: 0 SS K-major | 1 SS MN-major (lbo/sbo as chosen by `mn_swapped`) | 2 TS (A in TMEM)
// Each CTA issues `stages` x 12 MMAs (M=128, N=128, K=8), a commit per stage, waiting with a lag of 2 stages.
// writers != 0: warps 1..4 stream 16-byte shared-memory stores (64 KB per stage) meanwhile.
// kind 3: SS K-major SWIZZLE_128B tf32 | 4: SS K-major no-swizzle bf16 (kind::f16, K = 16) | 5: SS K-major tf32 N = 256
__global__ void __launch_bounds__(160) rate_probe(int kind, int mn_swapped, int stages, int writers, float* sink, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tslot;
  __shared__ volatile int done;
  uint8_t* sA = smem;                // 3 stage buffers x 32 KB (hi 16 KB + lo 16 KB) = 96 KB
  uint8_t* sB = smem + 96 * 1024;    // 128 KB weights (hi + lo)
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (224 * 1024) / 16; i += 160) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); done = 0; fence_barrier_init(); }
  if (warp == 0) umma::tmem_alloc<512>(&tslot);
  umma::fence_proxy_async_smem();
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem = tslot;
  bool issuer = (tid == 0);
  if (mn_swapped == 2) {   // warp-uniform branch + elect.sync (the CUTLASS idiom)
    issuer = false;
    if (warp == 0) {
      uint32_t pred = 0;
      asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
      issuer = pred != 0;
    }
  }
  if (issuer) {
    const uint32_t a0 = (uint32_t)__cvta_generic_to_shared(sA), b0 = (uint32_t)__cvta_generic_to_shared(sB);
    uint32_t idesc = kind == 1 ? idesc_tf32_major(128, 128, 1, 1) : umma::idesc_tf32(128, 128);
    if (kind == 4) idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (kind == 5) idesc = umma::idesc_tf32(128, 256);
    const long long t0 = clock64();
    for (int s = 0; s < stages; ++s) {
      const int slot = s % 3;
      if (s >= 3) mbar_wait(&bar[slot], ((s / 3) - 1) & 1);
      const uint32_t ah = a0 + slot * 32768, al = ah + 16384;
      const int kc = s & 3;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint64_t adh, adl, bdh, bdl;
        if (kind == 1) {
          const uint32_t kb = 4096u, lbo = mn_swapped == 1 ? 128u : kb, sbo = mn_swapped == 1 ? kb : 128u;
          adh = umma::smem_desc(ah + ks * kb, lbo, sbo); adl = umma::smem_desc(al + ks * kb, lbo, sbo);
          bdh = umma::smem_desc(b0 + (kc * 4 + ks) * kb, lbo, sbo); bdl = umma::smem_desc(b0 + 65536 + (kc * 4 + ks) * kb, lbo, sbo);
        } else {
          adh = umma::smem_desc(ah + ks * 256, 128, 1024); adl = umma::smem_desc(al + ks * 256, 128, 1024);
          const uint32_t sbo_b = kind == 5 ? 256u : 4096u;   // N = 256: overlapping row groups keep the reads inside the buffer
          bdh = umma::smem_desc(b0 + (kc * 4 + ks) * 256, 128, sbo_b); bdl = umma::smem_desc(b0 + 65536 + (kc * 4 + ks) * 256, 128, sbo_b);
        }
        const uint32_t d = tmem + (uint32_t)(s & 1) * (kind == 5 ? 256u : 128u);
        if (kind == 6) {   // independent accumulators for consecutive MMAs (is the N=128 cost a dependent-accumulate latency?)
          umma::mma_tf32_ss(tmem + 0, adl, bdh, idesc, 1);
          umma::mma_tf32_ss(tmem + 128, adh, bdl, idesc, 1);
          umma::mma_tf32_ss(tmem + 256, adh, bdh, idesc, 1);
          continue;
        }
        if (kind == 3) {
          // 128-byte swizzle: [128 rows x 32 k] atoms of 16 KB; row-group stride 1024 B; k-step = 32 B inside the row
          const uint64_t sah = with_swizzle128(umma::smem_desc(ah + ks * 32, 16, 1024)), sal = with_swizzle128(umma::smem_desc(al + ks * 32, 16, 1024));
          const uint64_t sbh = with_swizzle128(umma::smem_desc(b0 + kc * 16384 + ks * 32, 16, 1024));
          const uint64_t sbl = with_swizzle128(umma::smem_desc(b0 + 65536 + kc * 16384 + ks * 32, 16, 1024));
          umma::mma_tf32_ss(d, sal, sbh, idesc, 1);
          umma::mma_tf32_ss(d, sah, sbl, idesc, 1);
          umma::mma_tf32_ss(d, sah, sbh, idesc, 1);
        } else if (kind == 4) {
          mma_f16_ss(d, adl, bdh, idesc, 1);
          mma_f16_ss(d, adh, bdl, idesc, 1);
          mma_f16_ss(d, adh, bdh, idesc, 1);
        } else if (kind == 2) {
          mma_tf32_ts(d, tmem + 256 + ks * 8, bdh, idesc, 1);
          mma_tf32_ts(d, tmem + 256 + 32 + ks * 8, bdl, idesc, 1);
          mma_tf32_ts(d, tmem + 256 + ks * 8, bdl, idesc, 1);
        } else {
          umma::mma_tf32_ss(d, adl, bdh, idesc, 1);
          umma::mma_tf32_ss(d, adh, bdl, idesc, 1);
          umma::mma_tf32_ss(d, adh, bdh, idesc, 1);
        }
      }
      umma::commit(&bar[slot]);
    }
    // drain: the last commit covers every earlier MMA
    const int last = stages - 1;
    mbar_wait(&bar[last % 3], (last / 3) & 1);
    if (cycles) cycles[blockIdx.x] = clock64() - t0;
    done = 1;
  } else if (warp >= 1 && writers && tid >= 32) {
    // stream 16-byte stores into a scratch region past the operands (operands stay zero).
    // writers == 1: as fast as possible; writers == 2: 16 stores per warp per 768 cycles (the engine's
    // producer rate: 32 KB per 12-MMA stage).
    const uint32_t scratch = (uint32_t)__cvta_generic_to_shared(smem + 224 * 1024);
    const int t = tid - 32;
    int it = 0;
    long long next = clock64();
    while (!done) {
#pragma unroll
      for (int u = 0; u < 16; ++u)
        asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(scratch + (uint32_t)(t * 16)), "f"((float)(it + u)) : "memory");
      ++it;
      if (writers == 2) {
        next += 768;
        while (clock64() < next && !done) {}
      }
    }
    if (sink && it < 0) sink[0] = (float)it;
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<512>(tmem);
}

static double tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

int main() {
  std::vector<float> A((size_t)M * K), Bt((size_t)N * K), C((size_t)M * N);
  srand(1);
  for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& v : Bt) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  float *dA, *dB, *dC;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, Bt.size() * 4)); CK(cudaMalloc(&dC, C.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bt.data(), Bt.size() * 4, cudaMemcpyHostToDevice));
  const int smem = 3 * M * K * 4;
  CK(cudaFuncSetAttribute(gemm_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const char* names[5] = {"raw fp32 operands, 1 pass, K-major", "3-pass split, K-major", "3-pass, MN-major, desc.lbo = k-block stride, desc.sbo = mn-block stride",
                          "3-pass, MN-major, desc.lbo = mn-block stride, desc.sbo = k-block stride", "3-pass, A from TMEM, B K-major"};
  for (int mode = 0; mode < 5; ++mode) {
    CK(cudaMemset(dC, 0, C.size() * 4));
    gemm_probe<<<1, 128, smem>>>(dA, dB, dC, mode);
    CK(cudaGetLastError());
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d (%s): CUDA error %s\n", mode, names[mode], cudaGetErrorString(e)); return 1; }
    CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
    double e_exact = 0, e_trunc = 0, ref_norm = 0;
    for (int i = 0; i < M; ++i)
      for (int j = 0; j < N; ++j) {
        double ex = 0, tr = 0;
        for (int k = 0; k < K; ++k) {
          const float a = A[(size_t)i * K + k], b = Bt[(size_t)j * K + k];
          ex += (double)a * b; tr += tf32_trunc(a) * tf32_trunc(b);
        }
        const double c = C[(size_t)i * N + j];
        e_exact += (c - ex) * (c - ex); e_trunc += (c - tr) * (c - tr); ref_norm += ex * ex;
      }
    printf("mode %d (%s): rel L2 err vs exact %.3e, vs truncated-input product %.3e\n", mode, names[mode], sqrt(e_exact / ref_norm),
           sqrt(e_trunc / ref_norm));
  }

  // ---- throughput ----
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int rsmem = 224 * 1024 + 2048;
  CK(cudaFuncSetAttribute(rate_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, rsmem));
  float* sink; CK(cudaMalloc(&sink, 4));
  long long* dcyc; CK(cudaMalloc(&dcyc, sizeof(long long) * 256));
  std::vector<long long> hcyc(256);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int stages = 4096;
  struct Cfg { int kind, swapped, writers; const char* name; };
  Cfg cfgs[] = {{0, 0, 0, "SS K-major"}, {0, 0, 1, "SS K-major + smem writers"}, {1, 0, 0, "SS MN-major (cand.1)"},
                {0, 0, 2, "SS K-major + paced writers"}, {1, 1, 0, "SS MN-major (cand.2)"}, {2, 0, 0, "TS (A in TMEM)"}, {2, 0, 1, "TS + smem writers"},
                {0, 2, 0, "SS K-major, elect.sync issuer"}, {5, 2, 0, "SS tf32 N=256, elect.sync issuer"}, {6, 2, 0, "SS tf32 N=128, 3 rotating accumulators"}, {3, 0, 0, "SS K-major SWIZZLE_128B"}, {4, 0, 0, "SS bf16 (kind::f16) no swizzle"}, {5, 0, 0, "SS K-major tf32 N=256"}};
  for (const Cfg& c : cfgs) {
    rate_probe<<<sms, 160, rsmem>>>(c.kind, c.swapped, 64, c.writers, sink, nullptr);   // warm-up
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    rate_probe<<<sms, 160, rsmem>>>(c.kind, c.swapped, stages, c.writers, sink, dcyc);
    CK(cudaEventRecord(e1));
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("rate %s: CUDA error %s\n", c.name, cudaGetErrorString(e)); return 1; }
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaMemcpy(hcyc.data(), dcyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    double cyc = 0; for (int i = 0; i < sms; ++i) cyc += (double)hcyc[i]; cyc /= sms;
    const double mma_flop = (double)sms * stages * 12.0 * 2.0 * 128 * (c.kind == 5 ? 256 : 128) * (c.kind == 4 ? 16 : 8);
    printf("rate %-28s: %.3f ms, %.1f TFLOP/s tf32 MMA (= %.1f TFLOP/s algorithmic at 3 passes), %.1f SM cycles/MMA (clock64), implied clock %.0f MHz\n", c.name, ms,
           mma_flop / ms * 1e-9, mma_flop / 3 / ms * 1e-9, cyc / (stages * 12.0), cyc / (ms * 1e-3) * 1e-6);
  }
  return 0;
}
