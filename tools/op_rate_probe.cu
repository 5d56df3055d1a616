// Issue rate of the instructions the fused tcgen05 kernel's epilogue is made of (development probe): one CTA of 16 warps per SM,
// 8 independent chains per thread, cycles per warp-instruction and scheduler.  Also checks the bit pattern of
// cvt.rn.satfinite.tf32.f32 (low 13 bits zero, round to nearest even).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/op_rate_probe tools/op_rate_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int ITER = 2048;

template <int OP>
__device__ __forceinline__ void step(float (&f)[8], uint32_t (&u)[8], float2 (&p)[4]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if constexpr (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(1.0001f), "f"(0.5f));
    if constexpr (OP == 1) asm volatile("add.u32 %0, %0, 0x1000;" : "+r"(u[i]));
    if constexpr (OP == 2) asm volatile("and.b32 %0, %0, 0xFFFFE001;" : "+r"(u[i]));
    if constexpr (OP == 3) asm volatile("cvt.rn.satfinite.tf32.f32 %0, %1;" : "=r"(u[i]) : "f"(__uint_as_float(u[i])));
    if constexpr (OP == 4) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(__uint_as_float(u[i])), "f"(f[i]));
    if constexpr (OP == 5)
      asm volatile("{\n.reg .b16 lo, hi, m1;\nmov.b32 {lo, hi}, %1;\nmov.b16 m1, 0xBF80;\nfma.rn.f32.bf16 %0, hi, m1, %0;\n}\n" : "+f"(f[i]) : "r"(u[i]));
    if constexpr (OP == 6)
      asm volatile("{\n.reg .b16 lo, hi;\nmov.b32 {lo, hi}, %1;\nadd.rn.f32.bf16 %0, hi, %0;\n}\n" : "+f"(f[i]) : "r"(u[i]));
    if constexpr (OP == 7) asm volatile("shl.b32 %0, %0, 1;" : "+r"(u[i]));
    if constexpr (OP == 8) asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
    if constexpr (OP == 11) asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(u[i]));
    if constexpr (OP == 12) asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %2, 0;\nselp.f32 %0, %0, %1, q;\n}\n" : "+f"(f[i]) : "f"(f[(i + 1) & 7]), "r"(u[0] & 1));
    if constexpr (OP == 13) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
  }
  if constexpr (OP == 9) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      asm volatile("{\n.reg .b64 a, b, c;\nmov.b64 a, {%0, %1};\nmov.b64 b, {%2, %2};\nmov.b64 c, {%3, %3};\nfma.rn.f32x2 a, a, b, c;\nmov.b64 {%0, %1}, a;\n}\n"
                   : "+f"(p[i].x), "+f"(p[i].y) : "f"(1.0001f), "f"(0.5f));
      asm volatile("{\n.reg .b64 a, b, c;\nmov.b64 a, {%0, %1};\nmov.b64 b, {%2, %2};\nmov.b64 c, {%3, %3};\nfma.rn.f32x2 a, a, b, c;\nmov.b64 {%0, %1}, a;\n}\n"
                   : "+f"(f[2 * i]), "+f"(f[2 * i + 1]) : "f"(1.0001f), "f"(0.5f));
    }
  }
  if constexpr (OP == 10) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      asm volatile("{\n.reg .b64 a, b;\nmov.b64 a, {%0, %1};\nmov.b64 b, {%2, %2};\nmul.rn.f32x2 a, a, b;\nmov.b64 {%0, %1}, a;\n}\n"
                   : "+f"(p[i].x), "+f"(p[i].y) : "f"(1.0001f));
      asm volatile("{\n.reg .b64 a, b;\nmov.b64 a, {%0, %1};\nmov.b64 b, {%2, %2};\nmul.rn.f32x2 a, a, b;\nmov.b64 {%0, %1}, a;\n}\n"
                   : "+f"(f[2 * i]), "+f"(f[2 * i + 1]) : "f"(1.0001f));
    }
  }
}

template <int OP>
__global__ void __launch_bounds__(512, 1) rate(float* out, long long* cyc, float seed) {
  float f[8];
  uint32_t u[8];
  float2 p[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    f[i] = seed + 0.001f * (threadIdx.x + i);
    u[i] = __float_as_uint(f[i] * 1.37f);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = make_float2(f[i], f[i + 4]);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) step<OP>(f, u, p);
  const long long t1 = clock64();
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += f[i] + __uint_as_float(u[i]);
#pragma unroll
  for (int i = 0; i < 4; ++i) s += p[i].x + p[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void cvt_check(const float* x, uint32_t* y, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) asm("cvt.rn.satfinite.tf32.f32 %0, %1;" : "=r"(y[i]) : "f"(x[i]));
}

template <int OP>
static void run(const char* name, float* out, long long* cyc) {
  rate<OP><<<148, 512>>>(out, cyc, 1.0f);
  rate<OP><<<148, 512>>>(out, cyc, 1.0f);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < 148; ++i) mean += (double)h[i] / 148;
  // 16 warps per SM = 4 per scheduler, 8 instructions per step and thread
  printf("%-34s %6.2f cycles per warp-instruction and scheduler\n", name, mean / (ITER * 8.0 * 4.0));
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4);
  cudaMalloc(&cyc, 148 * 8);
  run<0>("FFMA", out, cyc);
  run<9>("FFMA2 (fma.rn.f32x2)", out, cyc);
  run<10>("FMUL2 (mul.rn.f32x2)", out, cyc);
  run<1>("IADD (add.u32 imm)", out, cyc);
  run<2>("LOP3 (and imm)", out, cyc);
  run<7>("SHF (shl)", out, cyc);
  run<8>("PRMT", out, cyc);
  run<12>("FSEL (selp.f32)", out, cyc);
  run<3>("F2FP.TF32 (cvt.rn.satfinite.tf32)", out, cyc);
  run<4>("F2FP.BF16 pack (cvt.rn.bf16x2)", out, cyc);
  run<5>("FHFMA.BF16 (fma.rn.f32.bf16)", out, cyc);
  run<6>("FHADD.BF16 (add.rn.f32.bf16)", out, cyc);
  run<11>("SHFL.BFLY", out, cyc);
  run<13>("MUFU.EX2", out, cyc);
  // bit pattern of the tf32 conversion
  const int n = 1 << 16;
  float* hx = (float*)malloc(n * 4);
  uint32_t* hy = (uint32_t*)malloc(n * 4);
  srand(1);
  for (int i = 0; i < n; ++i) {
    uint32_t b = ((uint32_t)rand() << 16) ^ (uint32_t)rand();
    b = (b & 0x807FFFFFu) | ((uint32_t)(100 + rand() % 50) << 23);
    if (i < 64) b = (b & 0xFFFFE000u) | 0x1000u;   // ties
    hx[i] = *(float*)&b;
  }
  float* dx;
  uint32_t* dy;
  cudaMalloc(&dx, n * 4);
  cudaMalloc(&dy, n * 4);
  cudaMemcpy(dx, hx, n * 4, cudaMemcpyHostToDevice);
  cvt_check<<<n / 256, 256>>>(dx, dy, n);
  cudaMemcpy(hy, dy, n * 4, cudaMemcpyDeviceToHost);
  int low_bits = 0, not_rne = 0, not_rna = 0;
  for (int i = 0; i < n; ++i) {
    const uint32_t b = *(uint32_t*)&hx[i];
    if (hy[i] & 0x1FFFu) ++low_bits;
    const uint32_t rna = (b + 0x1000u) & 0xFFFFE000u;
    const uint32_t rne = (b + 0xFFFu + ((b >> 13) & 1u)) & 0xFFFFE000u;
    if (hy[i] != rne) ++not_rne;
    if (hy[i] != rna) ++not_rna;
  }
  printf("cvt.rn.satfinite.tf32.f32 on %d values: %d with low bits set, %d differ from round-to-nearest-even, %d from round-half-away\n", n,
         low_bits, not_rne, not_rna);
  printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
