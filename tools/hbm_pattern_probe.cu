// What HBM bandwidth does the access pattern of tc_layer<fwd> get when nothing else happens?  148 persistent
// CTAs; per 120 KB tile 128 "producer" threads read the tile exactly as the layer kernel does (eight stages of
// 64-byte blocks, LDG.256) and 256 "epilogue" threads write a 120 KB tile as 16-byte chunks (30 STG.128 each).
// mode 0: readers and writers run freely (perfect overlap)   mode 1: the writers of tile i wait for its readers
// (the burst structure of the real kernel)                   mode 2: reads only   mode 3: writes only
// mode 4/5: as 2/0 with a per-lane prefetch.global.L2 of the same bytes one tile (gridDim tiles) ahead
// mode 6/7: as 2/0 with one cp.async.bulk.prefetch.L2 of the whole next tile per tile
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)
constexpr int NR = 240, H = 128;

__device__ __forceinline__ void ldg256(const void* p, float4& a, float4& b) {
  asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}

__global__ void __launch_bounds__(384) pattern(const float* __restrict__ in, float* __restrict__ out, int n_tiles, int mode, float* sink) {
  const int tid = threadIdx.x;
  float acc = 0.f;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    if (tid < 128) {
      if (mode >= 6 && tid == 0 && tile + gridDim.x < n_tiles)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(in + (size_t)(tile + gridDim.x) * NR * H), "r"(NR * H * 4) : "memory");
      if (mode == 8) {   // same bytes, fully coalesced: lane l of a warp reads chunk l of a 512-byte run
        for (int kc = 0; kc < 8; ++kc) {
          const float4* src = reinterpret_cast<const float4*>(in + (size_t)tile * NR * H + (size_t)kc * (NR * 16));
          float4 v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) { const int x = tid + 128 * i; if (x < NR * 4) v[i] = __ldg(src + x); else v[i] = make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
          for (int i = 0; i < 8; ++i) acc += v[i].x + v[i].w;
        }
      } else if (mode != 3) {
        const int kb = (tid >> 3) & 3, rb0 = (tid & 7) + 8 * (tid >> 5);
        for (int kc = 0; kc < 8; ++kc) {
          const unsigned char* src = reinterpret_cast<const unsigned char*>(in + (size_t)tile * NR * H) + (size_t)(kc * 2 + (kb >> 1)) * (NR * 32) + (kb & 1) * 64;
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int rb = rb0 + 32 * u;
            if (rb < NR / 4) {
              float4 a, b, c, d;
              if ((mode == 4 || mode == 5) && tile + gridDim.x < n_tiles)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(src + rb * 128 + (size_t)gridDim.x * NR * H * 4));
              ldg256(src + rb * 128, a, b);
              ldg256(src + rb * 128 + 32, c, d);
              acc += a.x + b.y + c.z + d.w;
            }
          }
        }
      }
      if (mode == 1) asm volatile("bar.arrive 1, 384;");
    } else {
      if (mode == 1) asm volatile("bar.sync 1, 384;");
      if (mode != 2 && mode != 4 && mode != 6 && mode != 8) {
        const int t = tid - 128, j = t & 127, half = t >> 7;
        float* io = out + (size_t)tile * NR * H + (size_t)(j >> 3) * (NR * 8) + (size_t)half * (5 * 32) + (size_t)(j & 7) * 4;
#pragma unroll
        for (int c = 0; c < 6; ++c)
#pragma unroll
          for (int g = 0; g < 5; ++g) *reinterpret_cast<float4*>(io + (c * 10 + g) * 32) = make_float4((float)tile, 1.f, 2.f, 3.f);
      }
    }
  }
  if (acc == 123.456f) sink[0] = acc;
}

int main() {
  const int n_tiles = 4320 * 4;
  float *in, *out, *sink;
  const size_t bytes = (size_t)n_tiles * NR * H * 4;
  CK(cudaMalloc(&in, bytes)); CK(cudaMalloc(&out, bytes)); CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(in, 0, bytes));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const char* names[9] = {"read + write, free running", "read then write per tile (burst)", "read only", "write only",
                          "read only + per-lane L2 prefetch", "read + write + per-lane L2 prefetch", "read only + bulk L2 prefetch", "read + write + bulk L2 prefetch", "read only, fully coalesced LDG.128"};
  for (int mode = 0; mode < 9; ++mode) {
    pattern<<<148, 384>>>(in, out, n_tiles, mode, sink);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    pattern<<<148, 384>>>(in, out, n_tiles, mode, sink);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double moved = ((mode <= 1 || mode == 5 || mode == 7) ? 2.0 : 1.0) * bytes;
    printf("%-36s: %.3f ms, %.2f TB/s\n", names[mode], ms, moved / ms * 1e-9);
  }
  return 0;
}
