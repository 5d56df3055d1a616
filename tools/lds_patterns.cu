// Shared-memory wavefront cost of the broadcast/partial-broadcast load patterns used by the jet GEMMs.
// Each pattern: every warp issues a long chain of independent LDS of the given width with a lane->address map;
// reported: cycles per LDS instruction per SM at saturation (8 warps/SM) ~= wavefronts per instruction.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

template <int W>  // W = 4, 8, 16 bytes
__global__ void k(const int* __restrict__ lane_off, float* out, long long* cyc, int iters) {
  extern __shared__ __align__(16) float sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = (float)i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const char* base = reinterpret_cast<const char*>(sm) + lane_off[lane];
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const char* p = base + ((u * 336 + it * 16) & 8191 & ~15);
      if (W == 16) { float4 v = *reinterpret_cast<const float4*>(p); acc0 += v.x; acc1 += v.y; acc2 += v.z; acc3 += v.w; }
      else if (W == 8) { float2 v = *reinterpret_cast<const float2*>(p); acc0 += v.x; acc1 += v.y; }
      else { acc0 += *reinterpret_cast<const float*>(p); }
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1 + acc2 + acc3;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  struct Pat { const char* name; int w; int (*f)(int); };
  Pat pats[] = {
    {"LDS.128 (lane&3)*16   [weights now: 4 distinct adjacent, 8-lane bcast]", 16, [](int l){return (l&3)*16;}},
    {"LDS.128 (lane&3)*32   [weights now, TC=8 stride]", 16, [](int l){return (l&3)*32;}},
    {"LDS.128 (lane&3)*336  [wgrad Z: 4 distinct, other banks]", 16, [](int l){return (l&3)*336;}},
    {"LDS.128 (lane>>2)*336 [wgrad A: 8 distinct by lr]", 16, [](int l){return (l>>2)*336;}},
    {"LDS.128 (lane>>3)*16  [4 distinct, uniform per quarter-warp]", 16, [](int l){return (l>>3)*16;}},
    {"LDS.128 (lane&7)*16   [8 distinct, each quarter has all 8]", 16, [](int l){return (l&7)*16;}},
    {"LDS.128 (lane>>2)*16  [8 distinct adjacent by lr]", 16, [](int l){return (l>>2)*16;}},
    {"LDS.128 uniform", 16, [](int l){return 0;}},
    {"LDS.128 lane*16 (all distinct)", 16, [](int l){return l*16;}},
    {"LDS.64  (lane>>2)*8   [a[c] now]", 8, [](int l){return (l>>2)*8;}},
    {"LDS.64  (lane&3)*8    [weights as 64-bit]", 8, [](int l){return (l&3)*8;}},
    {"LDS.64  (lane&3)*32", 8, [](int l){return (l&3)*32;}},
    {"LDS.64  lane*8 (all distinct)", 8, [](int l){return l*8;}},
    {"LDS.64  uniform", 8, [](int l){return 0;}},
    {"LDS.32  lane*4", 4, [](int l){return l*4;}},
    {"LDS.32  (lane&3)*4", 4, [](int l){return (l&3)*4;}},
    {"LDS.32  uniform", 4, [](int l){return 0;}},
  };
  int* d_off; float* d_out; long long* d_cyc;
  const int blocks = 148, threads = 256, iters = 2000;
  CK(cudaMalloc(&d_off, 32 * 4)); CK(cudaMalloc(&d_out, blocks * threads * 4)); CK(cudaMalloc(&d_cyc, blocks * 8));
  for (auto& p : pats) {
    int off[32]; for (int l = 0; l < 32; ++l) off[l] = p.f(l);
    CK(cudaMemcpy(d_off, off, sizeof(off), cudaMemcpyHostToDevice));
    for (int rep = 0; rep < 2; ++rep) {
      if (p.w == 16) k<16><<<blocks, threads, 32768>>>(d_off, d_out, d_cyc, iters);
      else if (p.w == 8) k<8><<<blocks, threads, 32768>>>(d_off, d_out, d_cyc, iters);
      else k<4><<<blocks, threads, 32768>>>(d_off, d_out, d_cyc, iters);
      CK(cudaDeviceSynchronize());
    }
    long long c[148]; CK(cudaMemcpy(c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += c[i]; avg /= blocks;
    // per SM: 8 warps x iters x 16 LDS instructions
    printf("%-75s %6.2f cycles/LDS/SM\n", p.name, avg / (8.0 * iters * 16));
  }
  return 0;
}
