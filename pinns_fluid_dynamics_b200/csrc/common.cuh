// Shared device-side descriptors and helpers for libpinnstep (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pinn {

constexpr int kChunk = 16;          // points per warp-chunk (8 row-lanes x 2 points)
constexpr int kMaxTerms = 8;        // PINN_MAX_TERMS_PER_SET
constexpr int kMaxOut = 4;          // PINN_MAX_OUT
constexpr int kMaxCh = 6;           // PINN_MAX_CH
constexpr int kMaxLaunchTerms = 64; // per-warp sum r^2 slots in shared memory

// One loss term as the kernels see it (host fills `scale` = 2*weight/(normalization*n_global)).
struct TermDev {
  float coef[kMaxOut][kMaxCh];
  float conv;
  int conv_k;
  float rhs_scale;
  float scale;          // 0 for test terms
  const float* rhs;     // [n_local] or nullptr
  int out_index;        // slot in the [T] tail of a workspace row
  int train;
};

// One point set (a "segment" of a launch): chunks [chunk_begin, chunk_begin + n_chunks).
struct SegDev {
  const float* pts;     // [n, D]
  float* y_out;         // optional [n, O] network values (pinn_forward), else nullptr
  long long n;
  int chunk_begin;
  int n_chunks;
  int n_terms;
  int pad_;
  TermDev terms[kMaxTerms];
};

__host__ __device__ constexpr int n_channels(int d, int order) {
  return order == 0 ? 1 : (order == 1 ? 1 + d : 3 + d);
}

// 1-D TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)),
               "r"(bytes));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(bar);
  while (!ok) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
               : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// tanh accurate to ~1 ulp over the whole range (libdevice tanhf: polynomial below 0.55, exp-based
// above); the loss tolerance (1e-5 relative) rules out tanh.approx.f32 (2^-11).
__device__ __forceinline__ float tanh_accurate(float z) { return tanhf(z); }

}  // namespace pinn
