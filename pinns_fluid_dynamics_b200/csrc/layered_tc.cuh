// Tensor-core layered engine for H = 128 tanh MLPs (BASELINE config 5: 3-128x8-3): the hidden-layer
// contractions run on tcgen05 (kind::tf32, FP32 accumulate in TMEM) with the 3-pass hi/lo split
//   x*w ~= x_lo*w_hi + x_hi*w_lo + x_hi*w_hi,   hi = rna_tf32(x), lo = rna_tf32(x - hi),
// because one TF32 pass misses the 1e-5 / 1e-4 parity tolerance (tools/tc_probe.cu).
//
// Measured facts this design rests on (tools/tc_probe.cu, tools/tc_layout_probe.cu on B200):
//   * an M=128,N=128,K=8 tf32 MMA costs ~110 cycles, an N=256 one 128 cycles (= the pipe's floor), so the
//     POINT ROWS go on the MMA N axis (240 or 256 rows per instruction) and the 128 neurons on M;
//   * tf32 operands must be K-major (MN-major descriptors read nothing); raw FP32 words are truncated by
//     the hardware, which is why both halves of the split are rounded to TF32 explicitly.
//
// Orientation: D[neuron j (TMEM lane)][row n (TMEM column)] = sum_k W[j][k] * Act[n][k].  A tile is P points
// x C channels, rows n = c*P + p, NR = C*P <= 240.  An epilogue thread owns one neuron of one TMEM lane
// quarter and sees every channel of a point in its own registers, so the tanh-jet (and its adjoint) needs no
// cross-lane traffic.
//
// HBM layout of every jet buffer (a-jets and z-bar alike), per tile: 16-byte chunks hold 4 CONSECUTIVE ROWS of
// one neuron, 8 neurons per 128-byte line,
//   off(n, k) = tile*NR*128 + (k/8)*(NR*8) + (n/4)*32 + (k%8)*4 + (n%4)          [floats]
// * epilogue threads (neuron k, 4-point groups) load / store whole chunks: a warp instruction touches 4 full
//   128-byte lines (the first version's 4-byte scattered accesses kept the LSU data pipe at 60-70 %);
// * it is the UMMA K-major core-matrix tiling of the TRANSPOSED operand, i.e. exactly what the weight-gradient
//   GEMM (contraction over rows) reads: its producer is a straight copy;
// * the forward / backward producers load 4-neuron x 4-row blocks (64 contiguous bytes) and transpose them in
//   registers; the row-group stride of their shared-memory stage blocks is padded to 528 B so those 16-byte
//   stores are bank-conflict free.
//
//   tc_prep_weights     hi/lo images of K_l (backward operand) and K_l^T (forward operand)
//   tc_layer1           a-jets of layer 1 (SIMT)
//   tc_layer<MODE=0>    a_l   = tanh-jet(a_{l-1} K_l + b_l)                       tcgen05, persistent
//   tc_out_layer        output GEMV, residuals, sum r^2, z-bar_L, K_out/b_out gradients (SIMT)
//   tc_wgrad            K_l-bar += a_{l-1}^T z-bar_l, b_l-bar                     tcgen05, persistent
//   tc_layer<MODE=1,2>  z-bar_{l-1} = tanh-jet-adjoint(z-bar_l K_l^T) in place    tcgen05, persistent
//   tc_layer1_grad      K_1-bar, b_1-bar from z-bar_1 (SIMT)
#pragma once
#include "common.cuh"
#include "fused_fp32.cuh"     // packed (f32x2) tanh-jet math: tanh2, jet2_from_a0, tanh_jet2_bwd
#include "layered_fp32.cuh"   // Jet, jet_fwd, jet_bwd
#include "umma.cuh"

namespace pinn {
namespace tc {

constexpr int kH = 128;

template <int D_, int ORDER_>
struct JetCfg {            // what the packed jet functions of fused_fp32.cuh need to know
  static constexpr int D = D_, ORDER = ORDER_, C = n_channels(D_, ORDER_), SX = D_ - 2, SY = D_ - 1;
};

template <int D, int ORDER>
struct Geo {
  static constexpr int C = n_channels(D, ORDER);
  static constexpr int P = C == 6 ? 40 : C == 5 ? 48 : C == 4 ? 56 : C == 3 ? 80 : 240;   // points per tile
  static constexpr int NR = C * P;                                                       // rows per tile = MMA N
  static_assert(NR % 16 == 0 && NR <= 240 && P % 8 == 0, "tile geometry");
};

// jet buffers: element (row n, neuron k) of a tile, in floats
template <int NR>
__host__ __device__ __forceinline__ size_t jet_off(int n, int k) {
  return (size_t)(k >> 3) * (NR * 8) + (size_t)(n >> 2) * 32 + (size_t)(k & 7) * 4 + (size_t)(n & 3);
}
// weight images: K-major core-matrix tiling of a 128 x 128 operand (row m, contraction index k), in floats
__host__ __device__ __forceinline__ size_t img_off(int m, int k) {
  return (size_t)(m >> 3) * (8 * kH) + (size_t)(k >> 2) * 32 + (size_t)(m & 7) * 4 + (size_t)(k & 3);
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_x4(uint32_t taddr, float (&v)[4]) {
  uint32_t r0, r1, r2, r3;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr) : "memory");
  v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
}
// 256-bit read-only global load (sm_100 LDG.E.256): one whole 32-byte sector per lane.  The producers read 64
// contiguous bytes per lane; as four LDG.128 every sector was requested twice (no L1 is left beside 226 KB of
// shared memory) -- measured 298 -> 246 us per forward layer with sector-clean loads.
__device__ __forceinline__ void ldg256(const void* p, float4& a, float4& b) {
  asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
// pull a contiguous global range into L2 ahead of use (no registers, no shared memory)
__device__ __forceinline__ void l2_prefetch_bulk(const void* gptr, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}
#ifndef PINN_TC_L2PF
#define PINN_TC_L2PF 4
#endif
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// round-to-nearest split: hi and lo are both exact TF32 values, so the tensor core's own truncation of its
// operands loses nothing and the residual x - hi - lo (<= 2^-22 |x|) has no sign bias.  (A truncating split
// leaves lo with 13 significant bits that the hardware cuts to 10: a one-sided 2^-21 error per product that
// adds up over 8 layers to 1.3e-5 on a loss term -- measured.)
__device__ __forceinline__ void tf32_split(float x, float& hi, float& lo) {
  hi = umma::rna_tf32(x);
  lo = umma::rna_tf32(x - hi);
}
// activation split on the hot path: hi rounded to nearest by the one-instruction conversion (SASS F2FP.SATFINITE.TF32.F32;
// cvt.rna.tf32 expands to an add and a mask on sm_100a), lo = x - hi exact and left for the tensor core to truncate.  lo's sign
// is independent of x's (hi is rounded, not truncated), so that truncation is zero-mean: no bias, error <= 2^-22 |x|.
__device__ __forceinline__ void tf32_split_fast(float x, float& hi, float& lo) {
  uint32_t h;
  asm("cvt.rn.satfinite.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  lo = x - hi;
}
__device__ __forceinline__ void tf32_split4(const float4& v, float4& hi, float4& lo) {
  tf32_split_fast(v.x, hi.x, lo.x); tf32_split_fast(v.y, hi.y, lo.y); tf32_split_fast(v.z, hi.z, lo.z); tf32_split_fast(v.w, hi.w, lo.w);
}

// ---- weight images -------------------------------------------------------------------------------
// per hidden->hidden layer: [fwd hi | fwd lo | bwd hi | bwd lo], each a 128x128 K-major core-matrix tiling.
//   fwd (A operand of  a K_l):     row m = output neuron j, k = input neuron i, value K_l[i][j]
//   bwd (A operand of z-bar K_l^T): row m = input neuron i,  k = output neuron j, value K_l[i][j]
constexpr int kImgFloats = kH * kH;
constexpr int kLayerImgFloats = 4 * kImgFloats;

__global__ void tc_prep_weights(const float* __restrict__ params, int off0, int stride, float* __restrict__ img) {
  const int l = blockIdx.y;
  const float* K = params + off0 + (size_t)l * stride;
  float* out = img + (size_t)l * kLayerImgFloats;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < kH * kH; idx += gridDim.x * blockDim.x) {
    const int i = idx / kH, j = idx % kH;
    float hi, lo;
    tf32_split(K[idx], hi, lo);
    const size_t f = img_off(j, i), b = img_off(i, j);
    out[f] = hi;
    out[kImgFloats + f] = lo;
    out[2 * kImgFloats + b] = hi;
    out[3 * kImgFloats + b] = lo;
  }
}

// ---- layer 1 (SIMT): thread = (neuron, 4-point group) -> one 16-byte chunk per channel ---------------------
template <int D, int ORDER>
__global__ void __launch_bounds__(256) tc_layer1(const float* __restrict__ params, const SegDev* __restrict__ segs,
                                                 const TileDev* __restrict__ tiles, float* __restrict__ act1) {
  using G = Geo<D, ORDER>;
  constexpr int C = G::C, P = G::P, NR = G::NR;
  const int j = threadIdx.x & (kH - 1), gj = threadIdx.x >> 7;
  const long long tile = blockIdx.x;
  const SegDev* __restrict__ seg = segs + tiles[tile].seg;
  const float* __restrict__ pts = seg->pts;
  const long long n = seg->n, p_begin = tiles[tile].p_begin;
  float zd[D];
#pragma unroll
  for (int i = 0; i < D; ++i) zd[i] = __ldg(params + i * kH + j);
  const float b = __ldg(params + D * kH + j);
  float* out = act1 + (size_t)tile * NR * kH + (size_t)(j >> 3) * (NR * 8) + (size_t)(j & 7) * 4;
  for (int g = gj; g < P / 4; g += 2) {
    float a[4][C];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      long long gp = p_begin + 4 * g + i;
      if (gp >= n) gp = n - 1;
      float z = b;
#pragma unroll
      for (int t = 0; t < D; ++t) z = fmaf(__ldg(pts + gp * D + t), zd[t], z);
      layered::jet_fwd<D, ORDER>(tanh_accurate(z), zd, 0.f, 0.f, a[i]);
    }
#pragma unroll
    for (int c = 0; c < C; ++c)
      *reinterpret_cast<float4*>(out + (size_t)(c * (P / 4) + g) * 32) = make_float4(a[0][c], a[1][c], a[2][c], a[3][c]);
  }
}

// ---- TMEM load helpers (no wait inside: issue several, then tmem_ld_wait) ---------------------------
template <int N>
__device__ __forceinline__ void tmem_ld_n(uint32_t taddr, float* v);
template <>
__device__ __forceinline__ void tmem_ld_n<4>(uint32_t taddr, float* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(taddr) : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld_n<8>(uint32_t taddr, float* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(taddr) : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld_n<16>(uint32_t taddr, float* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]),
                 "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]) : "r"(taddr) : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld_n<32>(uint32_t taddr, float* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]),
        "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]), "=f"(v[17]), "=f"(v[18]),
        "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]), "=f"(v[25]), "=f"(v[26]), "=f"(v[27]),
        "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
      : "r"(taddr) : "memory");
}
// W consecutive columns starting at taddr, as the largest power-of-two loads (W % 4 == 0)
template <int W, int O = 0>
__device__ __forceinline__ void tmem_ld_span(uint32_t taddr, float* v) {
  if constexpr (W - O >= 32) { tmem_ld_n<32>(taddr + O, v + O); tmem_ld_span<W, O + 32>(taddr, v); }
  else if constexpr (W - O >= 16) { tmem_ld_n<16>(taddr + O, v + O); tmem_ld_span<W, O + 16>(taddr, v); }
  else if constexpr (W - O >= 8) { tmem_ld_n<8>(taddr + O, v + O); tmem_ld_span<W, O + 8>(taddr, v); }
  else if constexpr (W - O >= 4) { tmem_ld_n<4>(taddr + O, v + O); tmem_ld_span<W, O + 4>(taddr, v); }
}
template <int REGS> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS)); }
template <int REGS> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS)); }

// ---- tcgen05 hidden-layer kernel -------------------------------------------------------------------
// The tensor core adds every MMA into its FP32 accumulator with TRUNCATION (measured, tools/
// tc_accum_probe.cu: -1.1e-6 relative bias after the 48 MMAs of a K = 128 contraction, 5x the error of an
// FFMA chain and one-sided, so it does not average out over layers).  The accumulator therefore only
// lives for one K-chunk of 32 (12 MMAs, small lo-terms first): two TMEM buffers ping-pong by chunk and the
// epilogue warps add each finished chunk into FP32 registers with round-to-nearest.  Warp roles (4 warp
// groups, registers redistributed with setmaxnreg):
//   WG0  warps 0-3    producers  global jets -> hi/lo split -> K-major stage blocks (2 stages prefetched in registers)
//   WG1-2 warps 4-11  epilogue   thread = neuron of lane quarter warp%4, half of the tile's points per group
//   WG3  warp 12      MMA issuer (one lane); warps 13-15 idle
// -DPINN_TC_PROFILE: CTA 0 accumulates clock64 intervals per warp role (tools/tc_role_profile.py)
#ifdef PINN_TC_PROFILE
__device__ unsigned long long g_tc_prof[32];
#define TCP_T0() const long long tcp_t0 = clock64()
#define TCP_ADD(slot) do { if (blockIdx.x == 0 && lane == 0) atomicAdd(&g_tc_prof[slot], (unsigned long long)(clock64() - tcp_t0)); } while (0)
#else
#define TCP_T0() do {} while (0)
#define TCP_ADD(slot) do {} while (0)
#endif
constexpr int kProdWarps = 4;
constexpr int kEpiWarps = 8;
constexpr int kLayerThreads = 512;
constexpr int kMmaWarp = 12;
constexpr int kStages = 3;
constexpr int kKC = 16;                       // K per pipeline stage (2 MMA k-steps)
constexpr int kStagesPerChunk = 2;            // accumulator lifetime: K = 32 (tools/tc_accuracy.py: 4 drains per tile keep the worst loss term at 4-6e-6; 2 drains 1.1e-5, none 2.3e-5)
constexpr int kChunks = kH / (kKC * kStagesPerChunk);

template <int NR>
struct LayerSmem {
  static constexpr int W_BYTES = 2 * kH * kH * 4;          // hi + lo images of the layer's A operand
  static constexpr int SBO = (kKC / 4) * 128 + 16;         // padded row-group stride of a stage block (528 B)
  static constexpr int HALF = (NR / 8) * SBO;              // one stage's hi (or lo) B block
  static constexpr int STAGE = 2 * HALF;
  static constexpr int BAR_OFF = W_BYTES + kStages * STAGE;
  static constexpr int TOTAL = BAR_OFF + 128;
};

// MODE 0: forward (bias + tanh-jet -> act_io written)
// MODE 1: backward hidden (act_io holds a_{l-1}, overwritten by z-bar_{l-1})
// MODE 2: backward into layer 1 (pre-activation jets are the constant rows of K1)
template <int D, int ORDER, int MODE>
__global__ void __launch_bounds__(kLayerThreads, 1) tc_layer(const float* __restrict__ w_img /* [hi|lo] */, const float* __restrict__ bias,
                                                             const float* __restrict__ act_in, float* __restrict__ act_io,
                                                             const float* __restrict__ params, int n_tiles) {
  using G = Geo<D, ORDER>;
  using S = LayerSmem<G::NR>;
  constexpr int C = G::C, P = G::P, NR = G::NR, PH = P / 2;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW = smem;
  uint8_t* sStage = smem + S::W_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);   // [kStages]
  uint64_t* empty = full + kStages;                                  // [kStages]
  uint64_t* tfull = empty + kStages;                                 // [2]
  uint64_t* tempty = tfull + 2;                                      // [2]
  uint64_t* wbar = tempty + 2;
  uint32_t* tslot = reinterpret_cast<uint32_t*>(wbar + 1);
  volatile int* prog = reinterpret_cast<volatile int*>(tslot + 1);   // tile index the producers are working on
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    *prog = 0;
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], kProdWarps); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], kEpiWarps); }
    mbar_init(wbar, 1);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) umma::tmem_alloc<512>(tslot);
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem = *tslot;

  if (warp < kProdWarps) {
    // ===== producers: two stages of loads in flight in registers (HBM latency x 21 B/cycle/SM ~ 30 KB) =====
    reg_dec<120>();
    if (tid == 0) {
      mbar_expect_tx(wbar, S::W_BYTES);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        tma_bulk_g2s(sW + i * (S::W_BYTES / 4), reinterpret_cast<const uint8_t*>(w_img) + i * (S::W_BYTES / 4), S::W_BYTES / 4, wbar);
    }
    constexpr int NKC = kH / kKC;
    const int my_tiles = n_tiles > (int)blockIdx.x ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int T = my_tiles * NKC;                          // this CTA's stage sequence: t -> (tile, k-chunk)
    int stage = 0;
    uint32_t phase = 0;
    // a stage = 16 neurons (2 groups of 8) x NR rows.  Block (rb, kb) = rows 4rb..4rb+3 x neurons 4kb..4kb+3 of
    // the stage = 64 contiguous bytes; thread t, pass u: kb = (t>>3)&3, rb = (t&7) + 8*(t>>5) + 32u, so the 8 lanes
    // of a 16-byte store phase hit 8 distinct bank groups (row-group stride 528 B).
    constexpr int NRB = NR / 4;
    const int kb = (tid >> 3) & 3, rb0 = (tid & 7) + 8 * (tid >> 5);
    auto issue = [&](int t, float4 (&v)[8]) {
      if constexpr (PINN_TC_L2PF > 0) {        // steady L2 prefetch PINN_TC_L2PF stages ahead of the register loads
        const int tp = t + PINN_TC_L2PF;
        if (tid < 2 && tp < T) {
          const int ptile = (int)blockIdx.x + (tp / NKC) * (int)gridDim.x, pkc = tp % NKC;
          l2_prefetch_bulk(reinterpret_cast<const uint8_t*>(act_in + (size_t)ptile * NR * kH) + (size_t)(pkc * 2 + tid) * (NR * 32), NR * 32);
          if constexpr (MODE == 1)
            l2_prefetch_bulk(reinterpret_cast<const uint8_t*>(act_io + (size_t)ptile * NR * kH) + (size_t)(pkc * 2 + tid) * (NR * 32), NR * 32);
        }
      }
      if (t >= T) return;
      const int tile = (int)blockIdx.x + (t / NKC) * (int)gridDim.x, kc = t % NKC;
      const uint8_t* src = reinterpret_cast<const uint8_t*>(act_in + (size_t)tile * NR * kH) +
                           (size_t)(kc * 2 + (kb >> 1)) * (NR * 32) + (kb & 1) * 64;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int rb = rb0 + 32 * u;
        if (rb < NRB) {
          const float4* p = reinterpret_cast<const float4*>(src + rb * 128);
          ldg256(p, v[4 * u], v[4 * u + 1]);                             // v[4u + r] = neuron 4kb + r, rows 4rb..4rb+3
          ldg256(p + 2, v[4 * u + 2], v[4 * u + 3]);
        }
      }
    };
    int t_cons = 0;
    auto consume = [&](const float4 (&v)[8]) {
      if (tid == 0 && t_cons % NKC == 0) *prog = t_cons / NKC;
      ++t_cons;
      { TCP_T0(); mbar_wait(&empty[stage], phase ^ 1u); if (warp == 0) TCP_ADD(0); }
      uint8_t* dst = sStage + stage * S::STAGE;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int rb = rb0 + 32 * u;
        if (rb < NRB) {
#pragma unroll
          for (int r = 0; r < 4; ++r) {                                   // row 4rb + r: the 4 neurons of this block
            const float e0 = r == 0 ? v[4 * u + 0].x : r == 1 ? v[4 * u + 0].y : r == 2 ? v[4 * u + 0].z : v[4 * u + 0].w;
            const float e1 = r == 0 ? v[4 * u + 1].x : r == 1 ? v[4 * u + 1].y : r == 2 ? v[4 * u + 1].z : v[4 * u + 1].w;
            const float e2 = r == 0 ? v[4 * u + 2].x : r == 1 ? v[4 * u + 2].y : r == 2 ? v[4 * u + 2].z : v[4 * u + 2].w;
            const float e3 = r == 0 ? v[4 * u + 3].x : r == 1 ? v[4 * u + 3].y : r == 2 ? v[4 * u + 3].z : v[4 * u + 3].w;
            float4 hi, lo;
            tf32_split4(make_float4(e0, e1, e2, e3), hi, lo);
            const int n = 4 * rb + r;
            const uint32_t off = (uint32_t)(n >> 3) * S::SBO + (uint32_t)kb * 128u + (uint32_t)(n & 7) * 16u;
            *reinterpret_cast<float4*>(dst + off) = hi;
            *reinterpret_cast<float4*>(dst + S::HALF + off) = lo;
          }
        }
      }
      umma::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[stage]);
      if (++stage == kStages) { stage = 0; phase ^= 1u; }
    };
    float4 va[8], vb[8], vc[8];
    TCP_T0();
    issue(0, va);
    issue(1, vb);
#pragma unroll 1
    for (int t = 0; t < T; t += 3) {
      issue(t + 2, vc);
      consume(va);
      if (t + 1 < T) { issue(t + 3, va); consume(vb); }
      if (t + 2 < T) { issue(t + 4, vb); consume(vc); }
    }
    if (warp == 0) TCP_ADD(1);
  } else if (warp >= kMmaWarp) {
    // ===== MMA issuer =====
    reg_dec<24>();
    if (warp == kMmaWarp && lane == 0) {
      mbar_wait(wbar, 0);
      const uint32_t idesc = umma::idesc_tf32(kH, NR);
      const uint32_t w_hi = (uint32_t)__cvta_generic_to_shared(sW), w_lo = w_hi + kH * kH * 4;
      const uint32_t st0 = (uint32_t)__cvta_generic_to_shared(sStage);
      int stage = 0, g = 0;
      uint32_t phase = 0;
      TCP_T0();
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#pragma unroll 1
        for (int ch = 0; ch < kChunks; ++ch, ++g) {
          const int buf = g & 1;
          { TCP_T0(); if (g >= 2) mbar_wait(&tempty[buf], (uint32_t)(((g >> 1) - 1) & 1)); TCP_ADD(2); }
          umma::fence_after_thread_sync();
          const uint32_t d = tmem + (uint32_t)buf * 256u;
          uint32_t acc = 0;
#pragma unroll 1
          for (int s = 0; s < kStagesPerChunk; ++s) {
            { TCP_T0(); mbar_wait(&full[stage], phase); TCP_ADD(3); }
            umma::fence_after_thread_sync();
            const uint32_t b_hi = st0 + stage * S::STAGE, b_lo = b_hi + S::HALF;
            const uint32_t ka = (uint32_t)((ch * kStagesPerChunk + s) * (kKC / 8)) * 256u;
            uint64_t a_hi_d[kKC / 8], a_lo_d[kKC / 8], b_hi_d[kKC / 8], b_lo_d[kKC / 8];
#pragma unroll
            for (int ks = 0; ks < kKC / 8; ++ks) {
              a_hi_d[ks] = umma::smem_desc(w_hi + ka + ks * 256, 128, 4096);
              a_lo_d[ks] = umma::smem_desc(w_lo + ka + ks * 256, 128, 4096);
              b_hi_d[ks] = umma::smem_desc(b_hi + ks * 256, 128, S::SBO);
              b_lo_d[ks] = umma::smem_desc(b_lo + ks * 256, 128, S::SBO);
            }
            // small terms first: they meet a small accumulator
#pragma unroll
            for (int ks = 0; ks < kKC / 8; ++ks) { umma::mma_tf32_ss(d, a_lo_d[ks], b_hi_d[ks], idesc, acc); acc = 1; }
#pragma unroll
            for (int ks = 0; ks < kKC / 8; ++ks) umma::mma_tf32_ss(d, a_hi_d[ks], b_lo_d[ks], idesc, 1);
#pragma unroll
            for (int ks = 0; ks < kKC / 8; ++ks) umma::mma_tf32_ss(d, a_hi_d[ks], b_hi_d[ks], idesc, 1);
            umma::commit(&empty[stage]);
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
          umma::commit(&tfull[buf]);
        }
      }
      TCP_ADD(4);
    }
  } else {
    // ===== epilogue: chunk partial sums TMEM -> FP32 registers; tanh-jet (or its adjoint) -> global =====
    // A thread owns neuron j and the points [half*PH, half*PH + PH) of the tile (PH = P/2, a multiple of 4): its
    // TMEM columns are C contiguous spans and its global data C*PH/4 whole 16-byte chunks at constant offsets.
    reg_inc<184>();
    const int q = warp & 3, half = (warp - kProdWarps) >> 2;
    const int j = q * 32 + lane;
    constexpr int NG = PH / 4;                 // 4-point groups per thread
    float bj = 0.f;
    float k1[D];
#pragma unroll
    for (int i = 0; i < D; ++i) k1[i] = 0.f;
    if constexpr (MODE == 0) bj = __ldg(bias + j);
    if constexpr (MODE == 2) {
#pragma unroll
      for (int i = 0; i < D; ++i) k1[i] = __ldg(params + i * kH + j);
    }
    const size_t thr_off = (size_t)(j >> 3) * (NR * 8) + (size_t)half * (NG * 32) + (size_t)(j & 7) * 4;
    int g = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      float* io = act_io + (size_t)tile * NR * kH + thr_off;       // chunk (c, g) = 4 points at io[(c*P/4 + g)*32]
      float acc[C][PH];
#pragma unroll 1
      for (int ch = 0; ch < kChunks; ++ch, ++g) {
        const int buf = g & 1;
        { TCP_T0(); mbar_wait(&tfull[buf], (uint32_t)((g >> 1) & 1)); if (warp == kProdWarps) TCP_ADD(5); }
        TCP_T0();
        umma::fence_after_thread_sync();
        const uint32_t tb = tmem + (uint32_t)buf * 256u + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * PH);
        if (ch == 0) {
#pragma unroll
          for (int c = 0; c < C; ++c) tmem_ld_span<PH>(tb + (uint32_t)(c * P), acc[c]);
          tmem_ld_wait();
        } else {
          constexpr int W0 = PH < 32 ? PH : 32;
#pragma unroll
          for (int c = 0; c < C; ++c) {
#pragma unroll
            for (int o = 0; o < PH; o += 32) {
              if (PH - o >= W0) {
                float t[W0];
                tmem_ld_span<W0>(tb + (uint32_t)(c * P + o), t);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < W0; ++i) acc[c][o + i] += t[i];
              } else {
                constexpr int W1 = PH % 32 == 0 ? 4 : PH % 32;    // tail of a span longer than 32 columns
                float u[W1];
                tmem_ld_span<W1>(tb + (uint32_t)(c * P + o), u);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < W1; ++i) acc[c][o + i] += u[i];
              }
            }
          }
        }
        umma::fence_before_thread_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[buf]);
        if (warp == kProdWarps) TCP_ADD(6);
      }
      // ---- elementwise: one 4-point group (one 16-byte chunk per channel) at a time.  (Keeping the next group's
      //      stored a-jets in flight was measured twice and is slower: the extra 24 registers spill.) ----
      TCP_T0();
#pragma unroll
      for (int gg = 0; gg < NG; ++gg) {
        float aj[C][4];
        if constexpr (MODE == 1) {
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const float4 t4 = *reinterpret_cast<const float4*>(io + (c * (P / 4) + gg) * 32);
            aj[c][0] = t4.x; aj[c][1] = t4.y; aj[c][2] = t4.z; aj[c][3] = t4.w;
          }
        } else if constexpr (MODE == 2) {
          const float4 t4 = *reinterpret_cast<const float4*>(io + gg * 32);
          aj[0][0] = t4.x; aj[0][1] = t4.y; aj[0][2] = t4.z; aj[0][3] = t4.w;
        }
        // two points per instruction (fma.rn.f32x2): the elementwise phase is a latency-bound dependent chain and
        // it is what keeps the TMEM buffers from being handed back to the MMA warp
        using JC = JetCfg<D, ORDER>;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int i0 = 4 * gg + 2 * h;
          float2 res[C];
          if constexpr (MODE == 0) {
            float2 zd[D], zxx = bc2(0.f), zyy = bc2(0.f);
#pragma unroll
            for (int t = 0; t < D; ++t)
              zd[t] = ORDER >= 1 ? make_float2(acc[(ORDER >= 1) ? 1 + t : 0][i0], acc[(ORDER >= 1) ? 1 + t : 0][i0 + 1]) : bc2(0.f);
            if constexpr (ORDER >= 2) {
              zxx = make_float2(acc[(ORDER >= 2) ? 1 + D : 0][i0], acc[(ORDER >= 2) ? 1 + D : 0][i0 + 1]);
              zyy = make_float2(acc[(ORDER >= 2) ? 2 + D : 0][i0], acc[(ORDER >= 2) ? 2 + D : 0][i0 + 1]);
            }
            jet2_from_a0<JC>(tanh2(add2(make_float2(acc[0][i0], acc[0][i0 + 1]), bc2(bj))), zd, zxx, zyy, res);
          } else {
            float2 a1[C], ab[C], zd1[D];
#pragma unroll
            for (int t = 0; t < D; ++t) zd1[t] = bc2(k1[t]);
#pragma unroll
            for (int c = 0; c < C; ++c) {
              a1[c] = (MODE == 2 && c > 0) ? bc2(0.f) : make_float2(aj[c][2 * h], aj[c][2 * h + 1]);
              ab[c] = make_float2(acc[c][i0], acc[c][i0 + 1]);
            }
            tanh_jet2_bwd<JC, MODE == 2>(a1, zd1, ab, res);
          }
#pragma unroll
          for (int c = 0; c < C; ++c) { acc[c][i0] = res[c].x; acc[c][i0 + 1] = res[c].y; }   // results replace the consumed sums
        }
#pragma unroll
        for (int c = 0; c < C; ++c)
          *reinterpret_cast<float4*>(io + (c * (P / 4) + gg) * 32) =
              make_float4(acc[c][4 * gg], acc[c][4 * gg + 1], acc[c][4 * gg + 2], acc[c][4 * gg + 3]);
      }
      if (warp == kProdWarps) TCP_ADD(7);
    }
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == kMmaWarp) umma::tmem_dealloc<512>(tmem);
}

// ---- tcgen05 weight gradient -----------------------------------------------------------------------
// gK[i][j] += sum_rows a_prev[row][i] * zbar[row][j]; gb[j] += sum over value-channel rows of zbar[row][j].
// The contraction runs over ROWS and the jet layout already is the K-major core-matrix tiling of the
// transposed operands (16-byte chunk = 4 rows of a neuron), so the producers copy 16-row slabs
// (16 neuron groups x 512 contiguous bytes) straight into the stage blocks, splitting hi/lo on the way.
// The TMEM accumulator lives for 8 stages (128 rows, 48 MMAs); drain warps add it into FP32 registers
// (same truncation argument as tc_layer) and flush the 128 x 128 block with atomics at the end.
constexpr int kWgThreads = 288;           // 4 producer warps, 4 drain warps, 1 MMA warp
constexpr int kWgMmaWarp = 8;
constexpr int kWgRows = 16;               // rows (K) per stage
constexpr int kWgStages = 6;
constexpr int kWgStagesPerChunk = 8;
constexpr int kWgOp = kH * kWgRows * 4;   // bytes of one operand block (128 neurons x 16 rows): 8192
constexpr int kWgStage = 4 * kWgOp;       // A hi, A lo, Z hi, Z lo
constexpr int kWgBarOff = kWgStages * kWgStage;
constexpr int kWgSmem = kWgBarOff + 128 + kH * 4;

// slabs: 16-row slabs of the batch; a tile holds NR/16 of them (NR % 16 == 0); value-channel rows are the
// first P rows of a tile = its first P/16 slabs when P % 16 == 0 -- in general rows [0, P): checked per chunk.
__global__ void __launch_bounds__(kWgThreads, 1) tc_wgrad(const float* __restrict__ act_prev, const float* __restrict__ zbar,
                                                          long long n_slabs, int NR, int P, float* __restrict__ rows,
                                                          size_t row_stride, size_t off_gk) {
  // this CTA's own workspace row (summed over the rows in a fixed order afterwards: no atomics, bit-reproducible)
  float* __restrict__ gK = rows + (size_t)blockIdx.x * row_stride + off_gk;
  float* __restrict__ gb = gK + (size_t)kH * kH;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kWgBarOff);
  uint64_t* empty = full + kWgStages;
  uint64_t* tfull = empty + kWgStages;   // [2]
  uint64_t* tempty = tfull + 2;          // [2]
  uint32_t* tslot = reinterpret_cast<uint32_t*>(tempty + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int slabs_per_tile = NR / kWgRows;
  const long long my_stages = n_slabs > (long long)blockIdx.x ? (n_slabs - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const long long my_chunks = (my_stages + kWgStagesPerChunk - 1) / kWgStagesPerChunk;
  if (tid == 0) {
    for (int s = 0; s < kWgStages; ++s) { mbar_init(&full[s], 4); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 4); }
    fence_barrier_init();
  }
  if (warp == kWgMmaWarp) umma::tmem_alloc<256>(tslot);
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem = *tslot;

  if (warp < 4) {
    // ===== producers: chunk x = tid + 128 u (u < 4) of a 512-chunk operand block: neuron group ig = x>>5,
    // row chunk rc = (x>>3)&3, neuron-in-group i7 = x&7; shared-memory offset = 16 x (the block is compact) =====
    float bsum[4] = {0.f, 0.f, 0.f, 0.f};
    const int rc = (tid >> 3) & 3, i7 = tid & 7, ig0 = tid >> 5;
    const size_t kb_bytes = (size_t)NR * 32;                     // bytes between neuron groups of a tile
    auto issue = [&](long long st, float4 (&v)[2][4], bool& val) {
      val = false;
      if (st >= n_slabs) return;
      const long long tile = st / slabs_per_tile;
      const int slab = (int)(st % slabs_per_tile);
      const int row0 = slab * kWgRows + rc * 4;                  // first of this thread's 4 rows inside the tile
      val = row0 < P;                                            // value-channel rows (P % 4 == 0)
      const size_t base = (size_t)tile * NR * kH * 4 + (size_t)(row0 >> 2) * 128 + i7 * 16;
#pragma unroll
      for (int mat = 0; mat < 2; ++mat) {
        const uint8_t* src = reinterpret_cast<const uint8_t*>(mat == 0 ? act_prev : zbar) + base;
#pragma unroll
        for (int u = 0; u < 4; ++u) v[mat][u] = *reinterpret_cast<const float4*>(src + (size_t)(ig0 + 4 * u) * kb_bytes);
      }
    };
    int stage = 0;
    uint32_t phase = 0;
    auto consume = [&](const float4 (&v)[2][4], bool val) {
      mbar_wait(&empty[stage], phase ^ 1u);
      uint8_t* dst = smem + stage * kWgStage;
#pragma unroll
      for (int mat = 0; mat < 2; ++mat)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float4 hi, lo;
          tf32_split4(v[mat][u], hi, lo);
          const uint32_t off = (uint32_t)(tid + 128 * u) * 16u;
          *reinterpret_cast<float4*>(dst + (mat * 2) * kWgOp + off) = hi;
          *reinterpret_cast<float4*>(dst + (mat * 2 + 1) * kWgOp + off) = lo;
          if (mat == 1 && val) bsum[u] += (v[mat][u].x + v[mat][u].y) + (v[mat][u].z + v[mat][u].w);
        }
      umma::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[stage]);
      if (++stage == kWgStages) { stage = 0; phase ^= 1u; }
    };
    // three slabs of loads in flight
    float4 va[2][4], vb[2][4], vc[2][4];
    bool ra, rb, rcv;
    const long long G = gridDim.x;
    issue(blockIdx.x, va, ra);
    issue(blockIdx.x + G, vb, rb);
#pragma unroll 1
    for (long long st = blockIdx.x; st < n_slabs; st += 3 * G) {
      issue(st + 2 * G, vc, rcv);
      consume(va, ra);
      if (st + G < n_slabs) { issue(st + 3 * G, va, ra); consume(vb, rb); }
      if (st + 2 * G < n_slabs) { issue(st + 4 * G, vb, rb); consume(vc, rcv); }
    }
    // bias gradient: thread (tid, u) always holds neuron 8*(ig0 + 4u) + i7; its four row chunks rc (lane bits 3, 4) are added
    // in a fixed order and the rc == 0 lane -- the only owner of that neuron in the CTA -- adds the sum to the CTA's row
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float b = bsum[u];
      b += __shfl_xor_sync(0xffffffffu, b, 8);
      b += __shfl_xor_sync(0xffffffffu, b, 16);
      if (rc == 0 && my_stages > 0) gb[8 * (ig0 + 4 * u) + i7] += b;
    }
  } else if (warp < 8) {
    // ===== drain warps: chunk partial sums -> FP32 registers -> global atomics =====
    const int q = warp & 3;
    const int i = q * 32 + lane;
    float acc[kH];
#pragma unroll
    for (int c = 0; c < kH; ++c) acc[c] = 0.f;
    for (long long ch = 0; ch < my_chunks; ++ch) {
      const int buf = (int)(ch & 1);
      mbar_wait(&tfull[buf], (uint32_t)((ch >> 1) & 1));
      umma::fence_after_thread_sync();
      const uint32_t tb = tmem + (uint32_t)buf * 128u + ((uint32_t)(q * 32) << 16);
#pragma unroll
      for (int c0 = 0; c0 < kH; c0 += 32) {
        float t[32];
        tmem_ld_n<32>(tb + c0, t);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c0 + c] += t[c];
      }
      umma::fence_before_thread_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);
    }
    if (my_chunks > 0) {
      float4* g4 = reinterpret_cast<float4*>(gK + (size_t)i * kH);       // row i of this CTA's K-bar block: one owner thread
#pragma unroll
      for (int c = 0; c < kH; c += 4) {
        float4 t = g4[c >> 2];
        t.x += acc[c]; t.y += acc[c + 1]; t.z += acc[c + 2]; t.w += acc[c + 3];
        g4[c >> 2] = t;
      }
    }
  } else if (warp == kWgMmaWarp && lane == 0) {
    const uint32_t idesc = umma::idesc_tf32(kH, kH);
    const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(smem);
    constexpr uint32_t SBO = (kWgRows / 4) * 128;              // 512: neuron-group stride inside an operand block
    int stage = 0;
    uint32_t phase = 0;
    long long done_stages = 0;
    for (long long ch = 0; ch < my_chunks; ++ch) {
      const int buf = (int)(ch & 1);
      if (ch >= 2) mbar_wait(&tempty[buf], (uint32_t)(((ch >> 1) - 1) & 1));
      umma::fence_after_thread_sync();
      const uint32_t d = tmem + (uint32_t)buf * 128u;
      uint32_t acc = 0;
      for (int s = 0; s < kWgStagesPerChunk && done_stages < my_stages; ++s, ++done_stages) {
        mbar_wait(&full[stage], phase);
        umma::fence_after_thread_sync();
        const uint32_t a_hi = s0 + stage * kWgStage, a_lo = a_hi + kWgOp, z_hi = a_lo + kWgOp, z_lo = z_hi + kWgOp;
#pragma unroll
        for (int ks = 0; ks < kWgRows / 8; ++ks) {
          umma::mma_tf32_ss(d, umma::smem_desc(a_lo + ks * 256, 128, SBO), umma::smem_desc(z_hi + ks * 256, 128, SBO), idesc, acc);
          acc = 1;
        }
#pragma unroll
        for (int ks = 0; ks < kWgRows / 8; ++ks)
          umma::mma_tf32_ss(d, umma::smem_desc(a_hi + ks * 256, 128, SBO), umma::smem_desc(z_lo + ks * 256, 128, SBO), idesc, 1);
#pragma unroll
        for (int ks = 0; ks < kWgRows / 8; ++ks)
          umma::mma_tf32_ss(d, umma::smem_desc(a_hi + ks * 256, 128, SBO), umma::smem_desc(z_hi + ks * 256, 128, SBO), idesc, 1);
        umma::commit(&empty[stage]);
        if (++stage == kWgStages) { stage = 0; phase ^= 1u; }
      }
      umma::commit(&tfull[buf]);
    }
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == kWgMmaWarp) umma::tmem_dealloc<256>(tmem);
}

// sum of v over the first `nact` lanes of a warp (the other lanes are not in the branch), fixed order; result in lane 0
__device__ __forceinline__ float warp_sum_active(float v, int lane, int nact) {
  const unsigned mask = nact >= 32 ? 0xffffffffu : ((1u << nact) - 1u);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float other = __shfl_down_sync(mask, v, o);
    if (lane + o < nact) v += other;
  }
  return v;
}

// ---- output layer + residuals + adjoint (SIMT), one tile per CTA iteration ---------------------------
// phase 1  thread = (point, k-slice): partial output jets J[c][o] over its slice of the 128 neurons (consecutive
//          lanes = consecutive points = consecutive 4-byte words of the 16-byte chunks), slices summed through
//          shared memory; one thread per point then forms the residuals, sum r^2 and the adjoint Jb[c][o].
// phase 2  thread = (neuron, half of the 4-point groups): a-bar = Jb . K_out^T, tanh-jet adjoint, z-bar_L stored
//          in place as whole 16-byte chunks; K_out gradients accumulate in registers (the neuron is fixed).
template <int D, int O, int ORDER, bool TRAIN>
__global__ void __launch_bounds__(256) tc_out_layer(const float* __restrict__ params, int off_ko, const SegDev* __restrict__ segs,
                                                    const TileDev* __restrict__ tiles, int n_tiles, float* __restrict__ actL,
                                                    float* __restrict__ rows, size_t row_stride, int n_params) {
  // this CTA's own workspace row: [gradient | term sums], summed over the rows in a fixed order afterwards
  float* __restrict__ grad = rows + (size_t)blockIdx.x * row_stride;
  float* __restrict__ sumsq = grad + n_params;
  using G = Geo<D, ORDER>;
  constexpr int C = G::C, P = G::P, NR = G::NR, H = kH, CO = C * O;
  constexpr int PW = (P + 31) / 32;                 // warps that hold points in the residual phase
  constexpr int SX = D - 2, SY = D - 1;
  constexpr int NS = 256 / P;                       // k-slices in phase 1
  constexpr int KS = (H + NS - 1) / NS;
  __shared__ float sKo[H * 4];
  __shared__ float sJp[NS * P * CO];
  __shared__ float sJb[P * CO];
  __shared__ float sSq[PW][kMaxTerms];              // per-warp partial sums, added in warp order
  __shared__ float sGbo[PW][kMaxOut];
  __shared__ float sGko[H * O];                     // K_out gradient partials of the h2 == 1 threads
  const int tid = threadIdx.x;
  const float* Ko = params + off_ko;
  const float* bo = Ko + H * O;
  for (int i = tid; i < H * 4; i += 256) sKo[i] = (i & 3) < O ? __ldg(Ko + (i >> 2) * O + (i & 3)) : 0.f;
  float gbo_acc = 0.f;                              // thread o < O: b_out gradient of this CTA over all its tiles
  __syncthreads();
  // phase-2 identity
  const int k2 = tid & (H - 1), h2 = tid >> 7;
  float ko2[O], gko[O];
#pragma unroll
  for (int o = 0; o < O; ++o) { ko2[o] = sKo[k2 * 4 + o]; gko[o] = 0.f; }
  // phase-1 identity
  const int p1 = tid % P, s1 = tid / P;
  const bool act1 = tid < NS * P;

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    float* base = actL + (size_t)tile * NR * H;
    const SegDev* __restrict__ seg = segs + tiles[tile].seg;
    const long long n = seg->n, p_begin = tiles[tile].p_begin;
    const int n_terms = seg->n_terms;
    // ---- phase 1 ----
    if (act1) {
      float J[C][O];
#pragma unroll
      for (int c = 0; c < C; ++c)
#pragma unroll
        for (int o = 0; o < O; ++o) J[c][o] = 0.f;
      const int k_end = (s1 + 1) * KS < H ? (s1 + 1) * KS : H;
      for (int k = s1 * KS; k < k_end; ++k) {
        const float4 kv = *reinterpret_cast<const float4*>(sKo + k * 4);
        const float kov[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float a = base[jet_off<NR>(c * P + p1, k)];
#pragma unroll
          for (int o = 0; o < O; ++o) J[c][o] = fmaf(a, kov[o], J[c][o]);
        }
      }
#pragma unroll
      for (int c = 0; c < C; ++c)
#pragma unroll
        for (int o = 0; o < O; ++o) sJp[(s1 * P + p1) * CO + c * O + o] = J[c][o];
    }
    __syncthreads();
    if (tid < P) {
      const int lane1 = tid & 31, warp1 = tid >> 5;
      const int nact1 = P - 32 * warp1 < 32 ? P - 32 * warp1 : 32;      // lanes of this warp that hold points
      const long long gp = p_begin + tid;
      const bool valid = gp < n;
      float J[C][O], Jb[C][O];
#pragma unroll
      for (int c = 0; c < C; ++c)
#pragma unroll
        for (int o = 0; o < O; ++o) {
          float v = c == 0 ? __ldg(bo + o) : 0.f;
#pragma unroll
          for (int s = 0; s < NS; ++s) v += sJp[(s * P + tid) * CO + c * O + o];
          J[c][o] = v;
          Jb[c][o] = 0.f;
        }
      if (seg->y_out != nullptr && valid) {
#pragma unroll
        for (int o = 0; o < O; ++o) seg->y_out[gp * O + o] = J[0][o];
      }
      float gb[O];
#pragma unroll
      for (int o = 0; o < O; ++o) gb[o] = 0.f;
#pragma unroll 1
      for (int t = 0; t < n_terms; ++t) {
        const TermDev* __restrict__ T = seg->terms + t;
        if (TRAIN && !T->train) continue;
        float r = 0.f;
#pragma unroll
        for (int o = 0; o < O; ++o)
#pragma unroll
          for (int c = 0; c < C; ++c) r = fmaf(__ldg(&T->coef[o][c]), J[c][o], r);
        float cv = 0.f;
        int ck = 0;
        if constexpr (ORDER >= 1 && O >= 2) {
          cv = __ldg(&T->conv);
          ck = __ldg(&T->conv_k);
          const float ukx = ck == 0 ? J[1 + SX][0] : J[1 + SX][1];
          const float uky = ck == 0 ? J[1 + SY][0] : J[1 + SY][1];
          r = fmaf(cv, fmaf(J[0][0], ukx, J[0][1] * uky), r);
        }
        if (T->rhs != nullptr && valid) r = fmaf(-__ldg(&T->rhs_scale), __ldg(T->rhs + gp), r);
        if (!valid) r = 0.f;
        {
          const float sq = warp_sum_active(r * r, lane1, nact1);
          if (lane1 == 0) sSq[warp1][t] = sq;
        }
        if constexpr (TRAIN) {
          const float rb = __ldg(&T->scale) * r;
#pragma unroll
          for (int o = 0; o < O; ++o)
#pragma unroll
            for (int c = 0; c < C; ++c) Jb[c][o] = fmaf(__ldg(&T->coef[o][c]), rb, Jb[c][o]);
          if constexpr (ORDER >= 1 && O >= 2) {
            const float ukx = ck == 0 ? J[1 + SX][0] : J[1 + SX][1];
            const float uky = ck == 0 ? J[1 + SY][0] : J[1 + SY][1];
            const float m = cv * rb;
            Jb[0][0] = fmaf(m, ukx, Jb[0][0]);
            Jb[0][1] = fmaf(m, uky, Jb[0][1]);
            const float m0 = ck == 0 ? m : 0.f, m1 = ck == 0 ? 0.f : m;
            Jb[1 + SX][0] = fmaf(m0, J[0][0], Jb[1 + SX][0]);
            Jb[1 + SY][0] = fmaf(m0, J[0][1], Jb[1 + SY][0]);
            Jb[1 + SX][1] = fmaf(m1, J[0][0], Jb[1 + SX][1]);
            Jb[1 + SY][1] = fmaf(m1, J[0][1], Jb[1 + SY][1]);
          }
        }
      }
      if constexpr (TRAIN) {
#pragma unroll
        for (int c = 0; c < C; ++c)
#pragma unroll
          for (int o = 0; o < O; ++o) sJb[tid * CO + c * O + o] = Jb[c][o];
#pragma unroll
        for (int o = 0; o < O; ++o) {
          const float g = warp_sum_active(Jb[0][o], lane1, nact1);
          if (lane1 == 0) sGbo[warp1][o] = g;
        }
      }
    }
    __syncthreads();
    // this tile's sums of squares go to its own set's slots (tiles of several sets share the launch)
    if (tid < n_terms && (!TRAIN || seg->terms[tid].train)) {
      float sq = 0.f;
#pragma unroll
      for (int w = 0; w < PW; ++w) sq += sSq[w][tid];
      sumsq[seg->terms[tid].out_index] += sq;
    }
    if constexpr (TRAIN) {
      if (tid < O) {
#pragma unroll
        for (int w = 0; w < PW; ++w) gbo_acc += sGbo[w][tid];
      }
    }
    if constexpr (TRAIN) {
      // ---- phase 2 ----
      float* io = base + (size_t)(k2 >> 3) * (NR * 8) + (size_t)(k2 & 7) * 4;
      for (int g = h2; g < P / 4; g += 2) {
        float a[C][4];
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float4 t4 = *reinterpret_cast<const float4*>(io + (size_t)(c * (P / 4) + g) * 32);
          a[c][0] = t4.x; a[c][1] = t4.y; a[c][2] = t4.z; a[c][3] = t4.w;
        }
        float z[C][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float* jb = sJb + (4 * g + i) * CO;
          float aj[C], ab[C], zb[C], k1[D];
#pragma unroll
          for (int t = 0; t < D; ++t) k1[t] = 0.f;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            aj[c] = a[c][i];
            float b = 0.f;
#pragma unroll
            for (int o = 0; o < O; ++o) {
              const float w = jb[c * O + o];
              b = fmaf(w, ko2[o], b);
              gko[o] = fmaf(aj[c], w, gko[o]);
            }
            ab[c] = b;
          }
          layered::jet_bwd<D, ORDER, false>(aj, k1, ab, zb);
#pragma unroll
          for (int c = 0; c < C; ++c) z[c][i] = zb[c];
        }
#pragma unroll
        for (int c = 0; c < C; ++c)
          *reinterpret_cast<float4*>(io + (size_t)(c * (P / 4) + g) * 32) = make_float4(z[c][0], z[c][1], z[c][2], z[c][3]);
      }
    }
    __syncthreads();      // sJp / sJb are reused by the next tile
  }
  if constexpr (TRAIN) {
    // neuron k2 lives in two threads (h2 = 0, 1): the h2 == 1 partial goes through shared memory, h2 == 0 adds and owns the slot
    if (h2 == 1) {
#pragma unroll
      for (int o = 0; o < O; ++o) sGko[k2 * O + o] = gko[o];
    }
    __syncthreads();
    if (h2 == 0) {
#pragma unroll
      for (int o = 0; o < O; ++o) grad[off_ko + k2 * O + o] += gko[o] + sGko[k2 * O + o];
    }
    if (tid < O) grad[off_ko + H * O + tid] += gbo_acc;
  }
}

// ---- K1 / b1 gradients from z-bar_1 (SIMT): thread = (neuron, 4-point group), block-strided over tiles ------
template <int D, int ORDER>
__global__ void __launch_bounds__(256) tc_layer1_grad(const float* __restrict__ zbar1, const SegDev* __restrict__ segs,
                                                      const TileDev* __restrict__ tiles, int n_tiles, float* __restrict__ rows,
                                                      size_t row_stride) {
  float* __restrict__ grad = rows + (size_t)blockIdx.x * row_stride;     // this CTA's own workspace row
  using G = Geo<D, ORDER>;
  constexpr int P = G::P, NR = G::NR;
  __shared__ float sG[(1 + D) * kH];
  const int j = threadIdx.x & (kH - 1), gj = threadIdx.x >> 7;
  float gk[D], gbv = 0.f;
#pragma unroll
  for (int i = 0; i < D; ++i) gk[i] = 0.f;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const float* zt = zbar1 + (size_t)tile * NR * kH + (size_t)(j >> 3) * (NR * 8) + (size_t)(j & 7) * 4;
    const SegDev* __restrict__ seg = segs + tiles[tile].seg;
    const float* __restrict__ pts = seg->pts;
    const long long n = seg->n, p_begin = tiles[tile].p_begin;
    for (int g = gj; g < P / 4; g += 2) {
      const float4 z4 = *reinterpret_cast<const float4*>(zt + (size_t)g * 32);
      const float z0[4] = {z4.x, z4.y, z4.z, z4.w};
      gbv += (z0[0] + z0[1]) + (z0[2] + z0[3]);
#pragma unroll
      for (int t = 0; t < D; ++t) {
        float v = 0.f;
        if constexpr (ORDER >= 1) {
          const float4 d4 = *reinterpret_cast<const float4*>(zt + (size_t)((1 + t) * (P / 4) + g) * 32);
          v = (d4.x + d4.y) + (d4.z + d4.w);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          long long gp = p_begin + 4 * g + i;
          if (gp >= n) gp = n - 1;
          v = fmaf(__ldg(pts + gp * D + t), z0[i], v);
        }
        gk[t] += v;
      }
    }
  }
  // neuron j lives in two threads (gj = 0, 1): gj == 1 hands its partials over through shared memory, gj == 0 owns the slots
  if (gj == 1) {
    sG[D * kH + j] = gbv;
#pragma unroll
    for (int i = 0; i < D; ++i) sG[i * kH + j] = gk[i];
  }
  __syncthreads();
  if (gj == 0) {
    grad[D * kH + j] += gbv + sG[D * kH + j];                             // [K1 | b1] are contiguous
#pragma unroll
    for (int i = 0; i < D; ++i) grad[i * kH + j] += gk[i] + sG[i * kH + j];
  }
}

}  // namespace tc
}  // namespace pinn
