// Shared device-side descriptors and helpers for libpinnstep (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pinn {

constexpr int kChunk = 16;          // points per warp-chunk (8 row-lanes x 2 points)
constexpr int kMaxTerms = 8;        // PINN_MAX_TERMS_PER_SET
constexpr int kMaxOut = 4;          // PINN_MAX_OUT
constexpr int kMaxCh = 6;           // PINN_MAX_CH
constexpr int kMaxLaunchTerms = 32; // per-warp sum r^2 slots in shared memory (the scripts' tables have at most 20 terms)

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
  int kind;             // 0: sum r^2, adjoint scale*r;  1 (|mean|): sum r, adjoint scale*(*sign)
  int pad_;
  const float* sign;    // kind 1: +-1 written by the pre-pass of pinn_loss_and_grad
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

// One tile of the layered tensor-core engine: P consecutive points of ONE point set (tiles never straddle sets, so all
// sets of a derivative order share one pipeline run).
struct TileDev {
  int seg;              // index into the launch table's segments
  int pad_;
  long long p_begin;    // first point of the tile inside its set
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

// tanh to ~2 ulp over the whole range, BRANCH-FREE (libdevice tanhf takes a data-dependent branch at |x| = 0.55,
// which serialises both sides in a warp and stops the compiler from interleaving the tanh of neighbouring
// neurons).  |x| < 0.55: odd minimax polynomial x + x^3 q(x^2), relative error 7e-8 (fit in tools/ notes);
// otherwise 1 - 2/(exp(2|x|) + 1) with ex2.approx and a Newton-refined rcp.approx.  The loss tolerance (1e-5
// relative) rules out tanh.approx.f32 (2^-11).
__device__ __forceinline__ float tanh_accurate(float z) {
#ifdef PINN_TANH_LIBDEVICE
  return tanhf(z);
#else
  const float az = fabsf(z);
  const float u = z * z;
  float q = fmaf(u, -0.006615748628973961f, 0.021312739700078964f);
  q = fmaf(q, u, -0.053910065442323685f);
  q = fmaf(q, u, 0.13333117961883545f);
  q = fmaf(q, u, -0.3333333134651184f);
  const float small = fmaf(z * u, q, z);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(az * -2.885390081777927f));   // exp(-2|z|)
  const float d = 1.0f + e;
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  r = r * fmaf(-d, r, 2.0f);                         // one Newton step: r = 1/(1+e) to ~1 ulp
  const float big = copysignf(fmaf(-2.0f * e, r, 1.0f), z);
  return az < 0.55f ? small : big;
#endif
}

}  // namespace pinn
