// tcgen05 probes behind the tcgen05 version of the H = 32 fused kernel (round 2; run on a B200 through gpurun):
//   A  correctness of  D[128 rows][32] = A[128][32] * W[32][32]^T  with the A operand in TENSOR MEMORY (TS form,
//      hi/lo images written by tcgen05.st), W as K-major hi/lo images in shared memory, 3 passes x 4 k-steps, at
//      non-zero D / A column offsets (what the forward / input-adjoint GEMMs of the fused kernel issue)
//   B  correctness + D lane map of the weight-gradient form  D[k][j] = sum_r a[r][k] z[r][j]  as kind::f16 (bf16)
//      MMAs with BOTH operands MN-major (no swizzle) straight from the [row][neuron] images, M = 64 and M = 128, N = 32
//   C  issue rates (cycles per MMA, 148 CTAs x 1 issuing lane) of the shapes above
//   D  latency of one epilogue -> MMA -> epilogue hand-over (tcgen05.st, mbarrier, 15 MMAs, commit, tcgen05.ld)
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include "../pinns_fluid_dynamics_b200/csrc/common.cuh"
#include "../pinns_fluid_dynamics_b200/csrc/umma.cuh"
using namespace pinn;
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
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
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
               "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void split_rn(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
  lo = __uint_as_float(__float_as_uint(x - hi) + 0x1000u);
}

// ---- A: TS tf32, N = 32 ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) ts_check(const float* __restrict__ A, const float* __restrict__ W, float* __restrict__ D,
                                                uint32_t d_col, uint32_t a_col) {
  __shared__ __align__(1024) uint8_t sW[2 * 4096];   // hi, lo: (n,k) -> (n/8)*1024 + (k/4)*128 + (n%8)*16 + (k%4)*4
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 32 * 32; i += 128) {
    const int n = i / 32, k = i % 32;
    float hi, lo;
    split_rn(W[i], hi, lo);
    const uint32_t off = umma::tile_offset(n, k, 1024);
    *reinterpret_cast<float*>(sW + off) = hi;
    *reinterpret_cast<float*>(sW + 4096 + off) = lo;
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) umma::tmem_alloc<512>(&tslot);
  umma::fence_proxy_async_smem();
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem = tslot;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  {
    const int row = tid;
    for (int k0 = 0; k0 < 32; k0 += 8) {
      float h[8], l[8];
      for (int j = 0; j < 8; ++j) split_rn(A[row * 32 + k0 + j], h[j], l[j]);
      tmem_st8(tmem + lane_base + a_col + k0, h);
      tmem_st8(tmem + lane_base + a_col + 32 + k0, l);
    }
    tmem_wait_st();
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  if (tid == 0) {
    const uint32_t w0 = (uint32_t)__cvta_generic_to_shared(sW);
    const uint32_t idesc = umma::idesc_tf32(128, 32);
    uint32_t acc = 0;
    for (int ks = 0; ks < 4; ++ks) {
      const uint64_t bh = umma::smem_desc(w0 + ks * 256, 128, 1024), bl = umma::smem_desc(w0 + 4096 + ks * 256, 128, 1024);
      mma_tf32_ts(tmem + d_col, tmem + a_col + 32 + ks * 8, bh, idesc, acc);   // lo * hi
      mma_tf32_ts(tmem + d_col, tmem + a_col + ks * 8, bl, idesc, 1);          // hi * lo
      mma_tf32_ts(tmem + d_col, tmem + a_col + ks * 8, bh, idesc, 1);          // hi * hi
      acc = 1;
    }
    umma::commit(&bar);
  }
  mbar_wait(&bar, 0);
  umma::fence_after_thread_sync();
  float v[32];
  umma::tmem_ld_32x32(tmem + lane_base + d_col, v);
  for (int j = 0; j < 32; ++j) D[tid * 32 + j] = v[j];
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<512>(tmem);
}

// ---- B: bf16 MN-major weight-gradient form ---------------------------------------------------------------------
// images: (row r, neuron n) -> (r/8)*512 + (n/8)*128 + (r%8)*16 + (n%8)*2 bytes; R rows, 32 neurons
__global__ void __launch_bounds__(128) mn_check(const float* __restrict__ a, const float* __restrict__ z, float* __restrict__ D,
                                                int R, int M, int swap_fields) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int img = (R / 8) * 512 + 4096;      // padded: the M = 64 / 128 descriptors read past the 32 real neurons
  uint8_t* sa = smem;
  uint8_t* sz = smem + img;
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 2 * img / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  __syncthreads();
  for (int i = tid; i < R * 32; i += 128) {
    const int r = i / 32, n = i % 32;
    const uint32_t off = (uint32_t)(r >> 3) * 512u + (uint32_t)(n >> 3) * 128u + (uint32_t)(r & 7) * 16u + (uint32_t)(n & 7) * 2u;
    *reinterpret_cast<__nv_bfloat16*>(sa + off) = __float2bfloat16(a[i]);
    *reinterpret_cast<__nv_bfloat16*>(sz + off) = __float2bfloat16(z[i]);
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) umma::tmem_alloc<32>(&tslot);
  umma::fence_proxy_async_smem();
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem = tslot;
  if (tid == 0) {
    const uint32_t a0 = (uint32_t)__cvta_generic_to_shared(sa), z0 = (uint32_t)__cvta_generic_to_shared(sz);
    const uint32_t idesc = idesc_bf16(M, 32, 1, 1);
    const uint32_t lbo = swap_fields ? 128u : 512u, sbo = swap_fields ? 512u : 128u;   // canonical: LBO = k-block stride, SBO = mn-block stride
    for (int ks = 0; ks < R / 16; ++ks) {
      const uint64_t ad = umma::smem_desc(a0 + ks * 1024, lbo, sbo), bd = umma::smem_desc(z0 + ks * 1024, lbo, sbo);
      mma_f16_ss(tmem, ad, bd, idesc, ks > 0);
    }
    umma::commit(&bar);
  }
  mbar_wait(&bar, 0);
  umma::fence_after_thread_sync();
  float v[32];
  umma::tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16), v);
  for (int j = 0; j < 32; ++j) D[tid * 32 + j] = v[j];
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<32>(tmem);
}

// ---- C: issue rates --------------------------------------------------------------------------------------------
// kind 0 TS tf32 N=32 (60 MMAs / commit) | 1 SS tf32 N=32 | 2 TS tf32 N=64 | 3 SS bf16 MN-major M=64 N=32 K=16 (120 / commit)
//      4 SS bf16 MN-major M=128 N=32 | 5 mix: 60 x kind 0 + 120 x kind 3 per commit | 6 TS tf32 N=32, 15 MMAs / commit
//      7 TS tf32 N=16 | 8 TS tf32 N=128 | 9 TS tf32 N=256 | 10 SS bf16 K-major M=128 N=32 K=16 | 11 TS bf16 (A in TMEM) N=32 K=16
__global__ void __launch_bounds__(64) rate2(int kind, int stages, long long* cycles, long long* n_mma) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 * 1024) / 16; i += 64) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
  if (warp == 0) umma::tmem_alloc<512>(&tslot);
  umma::fence_proxy_async_smem();
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem = tslot;
  if (tid == 32) {
    const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t wB = s0;                 // weights, 8 KB
    const uint32_t aS = s0 + 16384;         // SS activations 16 KB hi + 16 KB lo
    const uint32_t gA = s0 + 65536, gZ = s0 + 65536 + 24576;   // bf16 MN-major images (640 rows: 40 KB ... use 20 KB windows)
    long long count = 0;
    const long long t0 = clock64();
    for (int s = 0; s < stages; ++s) {
      if (s >= 2) mbar_wait(&bar[s & 1], ((s >> 1) - 1) & 1);
      if (kind == 0 || kind == 5 || kind == 6 || kind == 7 || kind == 2 || kind == 8 || kind == 9 || kind == 1) {
        const int N = kind == 2 ? 64 : kind == 7 ? 16 : kind == 8 ? 128 : kind == 9 ? 256 : 32;
        const uint32_t idesc = umma::idesc_tf32(128, N);
        const int tiles = kind == 6 ? 5 : (160 / N > 0 ? 160 / N : 1);
        const int kss = kind == 6 ? 1 : 4;
        for (int c = 0; c < tiles; ++c)
          for (int ks = 0; ks < kss; ++ks) {
            const uint64_t bh = umma::smem_desc(wB + ks * 256, 128, 1024), bl = umma::smem_desc(wB + 4096 + ks * 256, 128, 1024);
            const uint32_t d = tmem + (uint32_t)((c * N) % 192);
            if (kind == 1) {
              const uint64_t ah = umma::smem_desc(aS + ks * 256, 128, 1024), al = umma::smem_desc(aS + 16384 + ks * 256, 128, 1024);
              umma::mma_tf32_ss(d, al, bh, idesc, 1);
              umma::mma_tf32_ss(d, ah, bl, idesc, 1);
              umma::mma_tf32_ss(d, ah, bh, idesc, 1);
            } else {
              const uint32_t ah = tmem + 256 + (uint32_t)((c % 4) * 32 + ks * 8), al = ah + 128;
              mma_tf32_ts(d, al, bh, idesc, 1);
              mma_tf32_ts(d, ah, bl, idesc, 1);
              mma_tf32_ts(d, ah, bh, idesc, 1);
            }
            count += 3;
          }
      }
      if (kind == 3 || kind == 4 || kind == 5) {
        const uint32_t idesc = idesc_bf16(kind == 4 ? 128 : 64, 32, 1, 1);
        for (int ks = 0; ks < 20; ++ks)           // 20 k-steps of 16 rows, revisited twice: 40 k-steps x 3 passes = 120 MMAs
          for (int p = 0; p < 6; ++p) {
            const uint64_t ad = umma::smem_desc(gA + ks * 1024, 512, 128), bd = umma::smem_desc(gZ + ks * 1024, 512, 128);
            mma_f16_ss(tmem + 480, ad, bd, idesc, 1);
            count += 1;
          }
      }
      if (kind >= 12 && kind <= 17) {
        // channel-innermost issue order: consecutive MMAs accumulate into DIFFERENT D tiles (independent chains)
        // 12: TS tf32 N32, 5 D tiles | 13: SS tf32 N32, 5 D tiles | 14: TS tf32 N32, 2 D tiles | 15: TS tf32 N64 (hi|lo stacked B), 5 tiles of 64 cols? (uses 320 cols: A at 384)
        // 16: TS bf16 K16 N32, 5 D tiles | 17: TS tf32 N32 5 tiles, passes innermost over tiles (ks, pass, c)
        const int N = kind == 15 ? 64 : 32;
        const int rot = kind == 14 ? 2 : 5;
        const uint32_t idesc = kind == 16 ? idesc_bf16(128, 32, 0, 0) : umma::idesc_tf32(128, N);
        const uint32_t abase = kind == 15 ? 384u : 256u;
        const int kss = kind == 16 ? 2 : 4, passes = kind == 16 ? 6 : 3;
        for (int ks = 0; ks < kss; ++ks)
          for (int p = 0; p < passes; ++p)
            for (int c = 0; c < rot; ++c) {
              const uint32_t d = tmem + (uint32_t)(c * N);
              const uint64_t bd = kind == 16 ? umma::smem_desc(wB + (p & 1) * 4096 + ks * 256, 128, 512)
                                             : umma::smem_desc(wB + (p & 1) * 4096 + ks * 256, 128, 1024);
              if (kind == 13) {
                umma::mma_tf32_ss(d, umma::smem_desc(aS + (p == 0 ? 16384 : 0) + ks * 256, 128, 1024), bd, idesc, 1);
              } else if (kind == 16) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
                    "r"(tmem + abase + c * 16 + ks * 8), "l"(bd), "r"(idesc), "r"(1u)
                    : "memory");
              } else {
                mma_tf32_ts(d, tmem + abase + (uint32_t)((c % 4) * 32 + ks * 8) + (kind == 15 ? 0u : (p == 0 ? 128u : 0u)), bd, idesc, 1);
              }
              count += 1;
            }
      }
      if (kind == 18 || kind == 19) {
        // weight-gradient form with stacked images: M = 64 = [b1 | b2] neurons, N = 64 = [c1 | c2], K = 16: one MMA per k-step
        // 18: one accumulator (64 cols) | 19: two accumulators alternating
        const uint32_t idesc = idesc_bf16(64, 64, 1, 1);
        for (int ks = 0; ks < 40; ++ks) {
          const uint64_t ad = umma::smem_desc(gA + (ks % 20) * 1024, 1024, 128), bd = umma::smem_desc(gZ + (ks % 20) * 1024, 1024, 128);
          mma_f16_ss(tmem + (kind == 19 ? (uint32_t)(ks & 1) * 64u : 0u), ad, bd, idesc, 1);
          count += 1;
        }
      }
      if (kind == 20 || kind == 21) {
        // collector reuse of the A operand: 20: per k-step (lo x hi), (hi x lo: fill), (hi x hi: lastuse) | 21: one A for all 12 MMAs of a D tile
        const uint32_t idesc = umma::idesc_tf32(128, 32);
        for (int c = 0; c < 5; ++c)
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t bh = umma::smem_desc(wB + ks * 256, 128, 1024), bl = umma::smem_desc(wB + 4096 + ks * 256, 128, 1024);
            const uint32_t d = tmem + (uint32_t)(c * 32);
            const uint32_t ah = tmem + 256 + (uint32_t)((c % 4) * 32 + (kind == 21 ? 0 : ks * 8)), al = ah + 128;
#define MMA_COLL(Q, D_, A_, B_) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32" Q " [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(D_), "r"(A_), "l"(B_), "r"(idesc), "r"(1u) : "memory")
            if (kind == 20) {
              MMA_COLL("", d, al, bh);
              MMA_COLL(".collector::a::fill", d, ah, bl);
              MMA_COLL(".collector::a::lastuse", d, ah, bh);
            } else {
              if (ks == 0) MMA_COLL(".collector::a::fill", d, ah, bl); else MMA_COLL(".collector::a::use", d, ah, bl);
              MMA_COLL(".collector::a::use", d, ah, bh);
              if (ks == 3) MMA_COLL(".collector::a::lastuse", d, ah, bh); else MMA_COLL(".collector::a::use", d, ah, bh);
            }
            count += 3;
          }
      }
      if (kind == 10 || kind == 11) {
        const uint32_t idesc = idesc_bf16(128, 32, 0, 0);
        for (int c = 0; c < 5; ++c)
          for (int ks = 0; ks < 2; ++ks)
            for (int p = 0; p < 6; ++p) {
              const uint64_t bd = umma::smem_desc(wB + ks * 256, 128, 512);
              if (kind == 10) mma_f16_ss(tmem + c * 32, umma::smem_desc(aS + ks * 256, 128, 512), bd, idesc, 1);
              else {
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem + c * 32),
                    "r"(tmem + 256 + c * 16 + ks * 8), "l"(bd), "r"(idesc), "r"(1u)
                    : "memory");
              }
              count += 1;
            }
      }
      umma::commit(&bar[s & 1]);
    }
    const int last = stages - 1;
    mbar_wait(&bar[last & 1], (last >> 1) & 1);
    if (stages >= 2) mbar_wait(&bar[(last - 1) & 1], ((last - 1) >> 1) & 1);
    cycles[blockIdx.x] = clock64() - t0;
    n_mma[blockIdx.x] = count;
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<512>(tmem);
}

// ---- C2: do tcgen05.mma (async, one issuing lane) and mma.sync (8 warps) run concurrently or share the tensor pipe? ----
__global__ void __launch_bounds__(288) conc(int stages, int hmma_iters, long long* cyc_t, long long* cyc_h, float* sink) {
  __shared__ __align__(1024) uint8_t sW[8192];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 8192 / 4; i += 288) reinterpret_cast<float*>(sW)[i] = 0.f;
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
  if (warp == 0) umma::tmem_alloc<512>(&tslot);
  umma::fence_proxy_async_smem();
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem = tslot;
  const long long t0 = clock64();
  if (warp < 8) {
    float d[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
    unsigned a[4] = {(unsigned)tid, 1u, 2u, 3u}, b[2] = {5u, 7u};
    for (int it = 0; it < hmma_iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][3];
    if (s == 12345.f) sink[0] = s;
    if (tid == 0) cyc_h[blockIdx.x] = clock64() - t0;
  } else if (tid == 256 && stages > 0) {
    const uint32_t wB = (uint32_t)__cvta_generic_to_shared(sW);
    const uint32_t idesc = umma::idesc_tf32(128, 32);
    for (int s = 0; s < stages; ++s) {
      if (s >= 2) mbar_wait(&bar[s & 1], ((s >> 1) - 1) & 1);
      for (int c = 0; c < 5; ++c)
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t bh = umma::smem_desc(wB + ks * 256, 128, 1024), bl = umma::smem_desc(wB + 4096 + ks * 256, 128, 1024);
          const uint32_t d = tmem + (uint32_t)(c * 32), ah = tmem + 256 + (uint32_t)((c % 4) * 32 + ks * 8), al = ah + 128;
          mma_tf32_ts(d, al, bh, idesc, 1);
          mma_tf32_ts(d, ah, bl, idesc, 1);
          mma_tf32_ts(d, ah, bh, idesc, 1);
        }
      umma::commit(&bar[s & 1]);
    }
    const int last = stages - 1;
    mbar_wait(&bar[last & 1], (last >> 1) & 1);
    if (stages >= 2) mbar_wait(&bar[(last - 1) & 1], ((last - 1) >> 1) & 1);
    cyc_t[blockIdx.x] = clock64() - t0;
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<512>(tmem);
}

// ---- D: hand-over latency ---------------------------------------------------------------------------------------
// warps 0..3 (one per lane quadrant): tcgen05.st of 8 columns x 2, wait::st, fence, arrive on `ready`;
// warp 4 lane 0: wait `ready`, n_mma TS MMAs (N = 32), commit to `done`; warps 0..3: wait `done`, tcgen05.ld x32 of D.
__global__ void __launch_bounds__(160) handover(int iters, int n_mma, long long* cycles) {
  __shared__ __align__(1024) uint8_t sW[8192];
  __shared__ uint64_t ready, done;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 8192 / 4; i += 160) reinterpret_cast<float*>(sW)[i] = 0.f;
  if (tid == 0) { mbar_init(&ready, 128); mbar_init(&done, 1); fence_barrier_init(); }
  if (warp == 0) umma::tmem_alloc<512>(&tslot);
  umma::fence_proxy_async_smem();
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem = tslot;
  const long long t0 = clock64();
  if (warp < 4) {
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
      float h[8];
      for (int j = 0; j < 8; ++j) h[j] = acc + (float)j;
      tmem_st8(tmem + lane_base + 256, h);
      tmem_st8(tmem + lane_base + 264, h);
      tmem_wait_st();
      umma::fence_before_thread_sync();
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(&ready)) : "memory");
      mbar_wait(&done, it & 1);
      umma::fence_after_thread_sync();
      float v[32];
      umma::tmem_ld_32x32(tmem + lane_base, v);
      acc = v[0] * 1e-30f;
    }
    if (acc == 123.f) cycles[1] = 0;
  } else if (tid == 128) {
    const uint32_t w0 = (uint32_t)__cvta_generic_to_shared(sW);
    const uint32_t idesc = umma::idesc_tf32(128, 32);
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&ready, it & 1);
      umma::fence_after_thread_sync();
      for (int i = 0; i < n_mma; ++i) mma_tf32_ts(tmem + (i % 5) * 32, tmem + 256 + (i & 1) * 8, umma::smem_desc(w0, 128, 1024), idesc, 1);
      umma::commit(&done);
    }
  }
  __syncthreads();
  if (tid == 0) cycles[0] = clock64() - t0;
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<512>(tmem);
}

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
  srand(3);
  auto rnd = [] { return (float)rand() / RAND_MAX * 2.f - 1.f; };
  // ---- A ----
  {
    std::vector<float> A(128 * 32), W(32 * 32), D(128 * 32);
    for (auto& v : A) v = rnd();
    for (auto& v : W) v = rnd();
    float *dA, *dW, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dW, W.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
    const uint32_t cfg[3][2] = {{0, 256}, {96, 320}, {128, 448}};
    for (auto& c : cfg) {
      CK(cudaMemset(dD, 0, D.size() * 4));
      ts_check<<<1, 128>>>(dA, dW, dD, c[0], c[1]);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("A: CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
      double err = 0, nrm = 0;
      for (int r = 0; r < 128; ++r)
        for (int n = 0; n < 32; ++n) {
          double ex = 0;
          for (int k = 0; k < 32; ++k) ex += (double)A[r * 32 + k] * W[n * 32 + k];
          err += (D[r * 32 + n] - ex) * (D[r * 32 + n] - ex); nrm += ex * ex;
        }
      printf("A: TS tf32 M=128 N=32 K=32 3-pass, D col %u, A col %u: rel L2 err %.3e\n", c[0], c[1], sqrt(err / nrm));
    }
  }
  // ---- B ----
  {
    const int R = 128;
    std::vector<float> a(R * 32), z(R * 32), D(128 * 32);
    for (auto& v : a) v = bf16_round(rnd());
    for (auto& v : z) v = bf16_round(rnd());
    float *da, *dz, *dD;
    CK(cudaMalloc(&da, a.size() * 4)); CK(cudaMalloc(&dz, z.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dz, z.data(), z.size() * 4, cudaMemcpyHostToDevice));
    const int smem = 2 * ((R / 8) * 512 + 4096);
    CK(cudaFuncSetAttribute(mn_check, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    std::vector<double> E(32 * 32);
    for (int k = 0; k < 32; ++k)
      for (int j = 0; j < 32; ++j) {
        double s = 0;
        for (int r = 0; r < R; ++r) s += (double)a[r * 32 + k] * z[r * 32 + j];
        E[k * 32 + j] = s;
      }
    for (int M : {64, 128})
      for (int sw = 0; sw < 2; ++sw) {
        CK(cudaMemset(dD, 0, D.size() * 4));
        mn_check<<<1, 128, smem>>>(da, dz, dD, R, M, sw);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("B: M=%d swap=%d CUDA error %s\n", M, sw, cudaGetErrorString(e)); return 1; }
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        // which TMEM lane holds row k of the result?
        printf("B: bf16 MN-major M=%d N=32 K=%d, desc fields %s: row -> lane map:", M, R, sw ? "swapped" : "canonical");
        double worst = 0;
        int found = 0;
        for (int k = 0; k < 32; ++k) {
          int best = -1; double be = 1e30;
          for (int L = 0; L < 128; ++L) {
            double err = 0, nrm = 0;
            for (int j = 0; j < 32; ++j) { err += (D[L * 32 + j] - E[k * 32 + j]) * (D[L * 32 + j] - E[k * 32 + j]); nrm += E[k * 32 + j] * E[k * 32 + j]; }
            const double r = sqrt(err / nrm);
            if (r < be) { be = r; best = L; }
          }
          if (be < 1e-3) { ++found; if (be > worst) worst = be; }
          if (k < 4 || k == 15 || k == 16 || k == 31) printf(" %d->%d(%.1e)", k, best, be);
        }
        printf("  | rows found %d/32, worst rel err %.2e\n", found, worst);
      }
  }
  // ---- C ----
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  {
    const int rsmem = 128 * 1024;
    CK(cudaFuncSetAttribute(rate2, cudaFuncAttributeMaxDynamicSharedMemorySize, rsmem));
    long long *dc, *dn;
    CK(cudaMalloc(&dc, 8 * 256)); CK(cudaMalloc(&dn, 8 * 256));
    std::vector<long long> hc(256), hn(256);
    const char* names[22] = {"TS tf32 M128 N32 K8 (60/commit)", "SS tf32 M128 N32 K8", "TS tf32 M128 N64", "SS bf16 MN-major M64 N32 K16 (120/commit)",
                             "SS bf16 MN-major M128 N32 K16", "mix 60 TS tf32 N32 + 120 bf16 MN M64", "TS tf32 N32, 15/commit", "TS tf32 M128 N16",
                             "TS tf32 M128 N128", "TS tf32 M128 N256", "SS bf16 K-major M128 N32 K16", "TS bf16 M128 N32 K16",
                             "TS tf32 N32, 5 rotating D tiles", "SS tf32 N32, 5 rotating D tiles", "TS tf32 N32, 2 rotating D tiles", "TS tf32 N64, 5 rotating D tiles",
                             "TS bf16 K16 N32, 5 rotating D tiles", "(unused)", "SS bf16 MN-major M64 N64 K16, 1 accumulator", "SS bf16 MN-major M64 N64 K16, 2 accumulators",
                             "TS tf32 N32, collector: hi filled once per k-step, reused for hi x hi", "TS tf32 N32, one A operand reused by all 12 MMAs of a D tile"};
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int kind = 0; kind < 22; ++kind) {
      if (kind == 17) continue;
      const int stages = 1000;
      rate2<<<sms, 64, rsmem>>>(kind, 16, dc, dn);
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      rate2<<<sms, 64, rsmem>>>(kind, stages, dc, dn);
      CK(cudaEventRecord(e1));
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("C: kind %d CUDA error %s\n", kind, cudaGetErrorString(e)); return 1; }
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      CK(cudaMemcpy(hc.data(), dc, 8 * sms, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hn.data(), dn, 8 * sms, cudaMemcpyDeviceToHost));
      double cyc = 0; for (int i = 0; i < sms; ++i) cyc += (double)hc[i]; cyc /= sms;
      printf("C: %-44s: %.3f ms, %lld MMAs/CTA, %.1f SM cycles/MMA, %.0f cycles/commit-stage, clock %.0f MHz\n", names[kind], ms, hn[0],
             cyc / (double)hn[0], cyc / stages, cyc / (ms * 1e-3) * 1e-6);
    }
  }
  // ---- C2 ----
  {
    long long *dt, *dh; float* sink;
    CK(cudaMalloc(&dt, 8 * 256)); CK(cudaMalloc(&dh, 8 * 256)); CK(cudaMalloc(&sink, 4));
    long long ht[256], hh[256];
    const int cfg[3][2] = {{400, 0}, {0, 6000}, {400, 6000}};
    for (auto& c : cfg) {
      CK(cudaMemset(dt, 0, 8 * 256)); CK(cudaMemset(dh, 0, 8 * 256));
      conc<<<sms, 288>>>(c[0], c[1], dt, dh, sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("C2: CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      CK(cudaMemcpy(ht, dt, 8 * sms, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hh, dh, 8 * sms, cudaMemcpyDeviceToHost));
      printf("C2: %d tcgen05 stages (60 TS tf32 N32 MMAs each) + %d x 8 HMMA.1688 per warp x 8 warps: tcgen05 role %lld cycles (%.1f / MMA), mma.sync role %lld cycles (%.2f / HMMA / SM)\n",
             c[0], c[1], ht[0], c[0] ? (double)ht[0] / (c[0] * 60.0) : 0.0, hh[0], c[1] ? (double)hh[0] / (c[1] * 64.0) : 0.0);
    }
  }
  // ---- D ----
  {
    long long* dc; CK(cudaMalloc(&dc, 64));
    long long hc[2];
    for (int n_mma : {0, 1, 15, 60}) {
      handover<<<sms, 160>>>(2000, n_mma, dc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("D: CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      CK(cudaMemcpy(hc, dc, 16, cudaMemcpyDeviceToHost));
      printf("D: hand-over round trip with %2d MMAs: %.0f cycles per iteration\n", n_mma, (double)hc[0] / 2000.0);
    }
  }
  return 0;
}
