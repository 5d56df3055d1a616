// One-shot SUM all-reduce of the step's [P + T] vector over NVLink peer memory (ranks of ONE node, world <= 8).
//
// The vector is 9 KB: an NCCL all-reduce of that size is pure latency (~25 us inside the step's CUDA graph).  Here every rank
// owns a receive block that its peers map through CUDA IPC; ONE kernel (one CTA) per rank
//   1. stores its vector into slot [parity][rank] of EVERY rank's block (peer stores over NVLink),
//   2. publishes the step number in flag [parity][rank] of every block (release, system scope),
//   3. waits until all `world` flags of its OWN block carry the step number (acquire),
//   4. sums the `world` slots in rank order into the caller's buffer -- the same order on every rank: bit-identical results.
// Two parities alternate: a peer can only reach step e + 2 (and overwrite parity e) after it has received this rank's
// contribution to step e + 1, which is sent after step e has been summed.  The wait is bounded by wall-clock time (2 minutes); a time-out raises a
// flag the host can read (pinn_p2p_status) instead of hanging the GPU.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pinn {
namespace p2p {

constexpr int kMaxWorld = 8;

struct Peers {
  float* data[kMaxWorld];        // block of rank r: [2][world][cap] floats
  uint32_t* flags[kMaxWorld];    // flags of rank r: [2][kMaxWorld]
};

constexpr unsigned long long kTimeoutNs = 120ull * 1000ull * 1000ull * 1000ull;   // 2 minutes
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(1024, 1)
allreduce_oneshot_kernel(Peers peers, int world, int rank, int64_t cap, float* __restrict__ buf, int count, uint32_t* __restrict__ epoch_dev,
                         int* __restrict__ status) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  const uint32_t epoch = *epoch_dev + 1u;
  const int par = (int)(epoch & 1u);
  __syncthreads();                                         // everybody has read the step number before thread 0 advances it
  // 1. my vector -> slot [par][rank] of every rank's block
  for (int r = 0; r < world; ++r) {
    float* dst = peers.data[r] + ((size_t)par * world + rank) * cap;
    for (int i = tid; i < count; i += nthr) dst[i] = buf[i];
  }
  __threadfence_system();
  __syncthreads();
  // 2. publish, 3. wait
  if (tid < world) {
    st_release_sys(peers.flags[tid] + par * kMaxWorld + rank, epoch);
    const uint32_t* mine = peers.flags[rank] + par * kMaxWorld + tid;
    uint32_t spins = 0;
    unsigned long long t0 = 0;
    while (ld_acquire_sys(mine) != epoch) {
      if ((++spins & 1023u) == 0) {                        // wall-clock bound: ranks may legitimately be seconds apart (host work
        const unsigned long long now = globaltimer_ns();   // between steps), but a peer that died must not hang the GPU for ever
        if (t0 == 0) t0 = now;
        else if (now - t0 > kTimeoutNs) {
          atomicAdd(status, 1);
          break;
        }
      }
    }
  }
  __syncthreads();
  // 4. sum in rank order
  const float* src = peers.data[rank] + (size_t)par * world * cap;
  for (int i = tid; i < count; i += nthr) {
    float s = 0.f;
    for (int r = 0; r < world; ++r) s += __ldcv(src + (size_t)r * cap + i);   // written by peers: bypass stale lines
    buf[i] = s;
  }
  if (tid == 0) *epoch_dev = epoch;
}

}  // namespace p2p
}  // namespace pinn
