// Fused loss-step kernel for the 2-32x3-O tanh MLP with its hidden-layer GEMMs on tcgen05 (engine "fused_tcgen05").
//
// Same contract as fused_step_kernel (fused_fp32.cuh): Taylor-mode jets forward, residuals, weighted mean-square terms and the
// hand-written reverse sweep to parameter gradients in ONE kernel, every activation on chip.  Replaces (reference, per epoch)
// 3 x model(x) + 14 inner tape.gradient sweeps + nisaba's outer tape.gradient over the collocation set
// (cavity_steady.py:159-188,212-214,242).
//
// Work decomposition (measurements behind every choice: profiles/tc_probes_r02.md)
//   * a CTA (one per SM, persistent) works on a TILE of 128 points at a time.  TMEM lane = point: the 16 epilogue warps
//     (4 per lane quadrant = 4 per scheduler) own one point per thread and 8 of the 32 neurons each -- two half-octets
//     {8g + 4v .. +3}, g = 2 hs + u, for warp w = 4 (2u + v) + quadrant and hs = 0, 1 -- so every channel of a (point, neuron)
//     pair (value, d/dx, d/dy, d2/dx2, d2/dy2) is in one thread, the tanh-jet math needs no cross-lane traffic, and the
//     neuron octets 0, 1 (k-steps 0, 1 of the next GEMM) are complete after the first half of an epilogue phase.
//     Warp 16 issues the MMAs (whole warp in the loop, one lane by elect.sync: bare back-to-back UTCHMMA, 18 cycles each);
//     warps 17-19 only fill its warp group (setmaxnreg moves their registers to the epilogue warps: 112 each).
//   * forward GEMM of a hidden layer: per channel c an M = 128 (points) x N = 32 (neurons) x K = 32 tile, kind::tf32 with the
//     3-pass split x = hi + lo (lo*W_hi + hi*W_lo + hi*W_hi), FP32 accumulators in TENSOR MEMORY (columns 32c..), the
//     activation operand ALSO in tensor memory (TS form: the epilogue threads write the hi / lo images of their own row with
//     tcgen05.st -- no shared-memory operand traffic; from shared memory the same MMA costs 41 cycles instead of 18), weights
//     as K-major hi / lo images in shared memory.  hi and lo are rounded with the one-instruction conversion
//     cvt.rn.satfinite.tf32.f32 (SASS F2FP.TF32).  60 MMAs per layer and tile; the k-steps are issued as soon as the epilogue
//     has produced the 8 neurons they contract over.
//   * weight gradient K-bar_l = a_{l-1}^T z-bar_l (contraction over the 128 x C rows of the tile): kind::f16 (bf16) MMAs with
//     BOTH operands MN-major straight from [row][neuron] images in shared memory, M = 64 = (b1 | b2 part) x 32 neurons,
//     N = 32, K = 16 rows per instruction.  The a-jets and z-bars are kept in shared memory ONLY as these images: a bf16
//     PAIR per value (b1 = bf16(x), b2 = bf16(x - b1): 16 mantissa bits; remainder by the mixed-precision FHFMA.BF16), 4 bytes
//     like the FP32 value they replace; the reverse sweep reads them back as b1 + b2 (they only feed gradients, tolerance
//     1e-4; the forward pass -- the loss values, tolerance 1e-5 -- never sees them: it goes registers -> tensor memory).
//     Image layout: every thread writes 16-byte atoms (STS.128, conflict-free) -- on the A side [b1 | b2] of the 4 neurons of
//     a half-octet, on the Z side one part of the 4 + 4 neurons it owns in the two halves of a phase; the accumulator's row /
//     column order follows and is undone in the drain.  The accumulator (tensor memory, 32 columns) is drained once per
//     tile into FP32 totals in shared memory (the tensor core truncates when it adds).
//   * input-adjoint GEMM a-bar_{l-1} = z-bar_l K_l^T: on the SAME bf16 pairs (kind::f16, operand in tensor memory: a 32-bit
//     column holds two neighbouring neurons; weights as a bf16 pair w1 + w2, K-major): b2 w1 + b1 w2 + b1 w1, K = 16 per
//     instruction, 30 MMAs per layer and tile -- gradients only, no operand split in the epilogue.
//   * residuals: the launch's loss terms sit in shared memory (24 words each); the four threads of a point share them -- thread h
//     evaluates the terms t = h (mod 4) as packed dot products, the adjoint seeds travel through 4 tensor-memory columns.
//   * small gradients (K1, b1, b_l, K_out, b_out): per-tile multi-value warp reductions into per-warp shared-memory
//     accumulators; at the end the CTA writes ONE workspace row, finalize_rows_kernel sums the rows in a fixed order
//     (no atomics anywhere: bit-reproducible).
//
// Tensor-memory map (512 columns, lane = point): D_c at 32c; forward operand A_c hi at 160 + 64c, lo at 160 + 64c + 32;
// adjoint operand b1 at 160 + 64c + [0, 16), b2 at 160 + 64c + 32 + [0, 16); output-jet exchange of warp h at 160 + 64h + [16, 32)
// (never written during the reverse sweep); weight-gradient accumulator at 480.
// Shared-memory map: TcCfg.
#pragma once
#include <cuda_bf16.h>

#include "fused_fp32.cuh"

namespace pinn {
namespace ftc {

template <int D_, int O_>
struct TcCfg {
  static constexpr int D = D_, H = 32, L = 3, O = O_, ORDER = 2;
  static constexpr int C = 3 + D;
  static constexpr int SX = D - 2, SY = D - 1;
  static constexpr int TP = 128;                    // points per tile
  static constexpr int NEPI = 16;                   // epilogue warps
  static constexpr int THREADS = 32 * (NEPI + 4);   // + the warp group of the MMA warp
  static constexpr int EPI_REGS = 112, AUX_REGS = 32;      // 16 x 32 x 112 + 4 x 32 x 32 = 640 x 96: setmaxnreg only redistributes the launch allocation
  static_assert(D == 2, "tensor-memory budget: 32C + 64C + 32 columns <= 512 needs C = 5");
  static constexpr uint32_t COL_D = 0, COL_A = 32 * C, COL_W = COL_A + 64 * C;
  static_assert(COL_W + 32 <= 512, "tensor memory exhausted");
  // shared memory (bytes)
  static constexpr int IMG_BYTES = C * TP * 32 * 4;               // [row = (c, p)][neuron] bf16 pairs: 81920
  static constexpr int OFF_X = 0, OFF_Y = IMG_BYTES;
  static constexpr int OFF_A1 = 2 * IMG_BYTES;                    // tanh(z1): a1[j][p] fp32
  static constexpr int OFF_W = OFF_A1 + 32 * TP * 4;              // forward weight images [(l-2)][hi|lo] of 4096 bytes (tf32)
  static constexpr int OFF_WB = OFF_W + 4 * 4096;                 // adjoint weight images [(l-2)][w1|w2] of 2048 bytes (bf16)
  static constexpr int TOT_LD = 36;                               // row stride of the totals: the drain's 8 rows x 16 bytes hit 32 distinct banks
  static constexpr int OFF_TOT = OFF_WB + 4 * 2048;               // weight-gradient totals [2][32][TOT_LD] fp32
  static constexpr int OFF_SMALL = OFF_TOT + 2 * 32 * TOT_LD * 4; // K1 [D][32] | b [3][32] | K_out [32][4] | b_out [4]
  static constexpr int S_K1 = 0, S_B = D * 32, S_KO = S_B + 3 * 32, S_BO = S_KO + 32 * 4, SMALL_FLOATS = S_BO + 4;
  static constexpr int OFF_SG = OFF_SMALL + SMALL_FLOATS * 4;     // per epilogue warp: small-gradient accumulators of its 8 neurons
  static constexpr int SG_K1 = 0, SG_B1 = D * 8, SG_B2 = SG_B1 + 8, SG_B3 = SG_B2 + 8, SG_KO = SG_B3 + 8, SG_BO = SG_KO + 32,
                       SG_FLOATS = SG_BO + 4;
  static constexpr int OFF_SSQ = OFF_SG + NEPI * SG_FLOATS * 4;   // warps 0..3: sum r^2 per term slot
  static constexpr int OFF_TERM = OFF_SSQ + 4 * kMaxLaunchTerms * 4;   // the launch's terms, TERM_WORDS words each
  static constexpr int TERM_WORDS = 24;
  static constexpr int T_CONV = 15, T_RHS_SCALE = 16, T_SCALE = 17, T_FLAGS = 18, T_OUT = 19, T_RHS = 20, T_SIGN = 22;
  static constexpr int OFF_SEGT = OFF_TERM + kMaxLaunchTerms * TERM_WORDS * 4;   // first staged term of each segment
  static constexpr int OFF_SEG = OFF_SEGT + kMaxLaunchTerms * 4;                 // per segment: chunk_begin | n_terms | n | pts | y_out (8 words)
  static constexpr int OFF_BAR = OFF_SEG + kMaxLaunchTerms * 32;
  static constexpr int SMEM_BYTES = OFF_BAR + 16 * 8;
  static_assert(SMEM_BYTES <= 232448, "shared memory exhausted");
  static_assert(O * C <= 15, "staged term layout");
  static constexpr int P = D * H + H + (L - 1) * (H * H + H) + H * O + O;
  static_assert(P * 4 <= IMG_BYTES, "parameter staging area");
  __host__ __device__ static constexpr int offK(int l) { return D * H + H + (l - 2) * (H * H + H); }   // l = 2..L
  static constexpr int OFF_KO = D * H + H + (L - 1) * (H * H + H);
  static constexpr int OFF_BO = OFF_KO + H * O;
};

// Development aid (-DPINN_TC_PROFILE, tools/tc_phase_profile.py): cycles per phase and warp, summed over the tiles of a CTA
#ifdef PINN_TC_PROFILE
__device__ unsigned long long g_tc_prof[160][20][16];
#define TC_PROF_DECL long long prof_t = clock64();
#define TC_PROF(k)                                                                                         \
  do {                                                                                                     \
    const long long now_ = clock64();                                                                      \
    if ((threadIdx.x & 31) == 0) atomicAdd(&g_tc_prof[blockIdx.x][threadIdx.x >> 5][k], (unsigned long long)(now_ - prof_t)); \
    prof_t = now_;                                                                                         \
  } while (0)
// raw time stamps of CTA 0, tiles 2..9 of its sequence: g_tc_trace[tile][warp][event] (tools/tc_trace.py)
__device__ long long g_tc_trace[8][20][24];
#define TC_TRACE(seq, ev)                                                                                  \
  do {                                                                                                     \
    if (blockIdx.x == 0 && (seq) >= 2 && (seq) < 10 && (threadIdx.x & 31) == 0)                            \
      g_tc_trace[(seq) - 2][threadIdx.x >> 5][ev] = clock64();                                             \
  } while (0)
// prologue / epilogue stamps of CTA 0, thread 0
__device__ long long g_tc_stage[16];
#define TC_STAGE(k)                                                          \
  do {                                                                       \
    if (blockIdx.x == 0 && threadIdx.x == 0) g_tc_stage[k] = clock64();      \
  } while (0)
#else
#define TC_PROF_DECL
#define TC_PROF(k)
#define TC_TRACE(seq, ev)
#define TC_STAGE(k)
#endif

// barriers (uint64 slots at OFF_BAR)
enum { B_AREADY = 0 /* ..3 */, B_DLOADED = 4, B_DFULL = 5, B_IMG = 6, B_WDONE = 7, B_STAGE = 8 };

// ---- PTX helpers ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
// barriers by their 32-bit shared-memory address (computed once: a generic-to-shared conversion per use costs an S2R)
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
template <int REGS> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS)); }
template <int REGS> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS)); }
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
               "r"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void mma_bf16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
               "r"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
// bf16 MMAs with both operands from shared memory; the A operand is kept in / taken from the collector buffer (the two MMAs of
// a k-step share it)
__device__ __forceinline__ void mma_bf16_ss_keep_a(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void mma_bf16_ss_reuse_a(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
// thread i of the warp: tensor-memory lane (quadrant base + i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]),
        "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
      : "r"(taddr)
      : "memory");
}
// 4 columns of each of the 5 accumulator tiles (stride 32 columns) in one go, one wait
__device__ __forceinline__ void tmem_ld4x5(uint32_t taddr, float (&v)[5][4]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%20];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%4, %5, %6, %7}, [%21];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%8, %9, %10, %11}, [%22];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%12, %13, %14, %15}, [%23];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%16, %17, %18, %19}, [%24];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=f"(v[0][0]), "=f"(v[0][1]), "=f"(v[0][2]), "=f"(v[0][3]), "=f"(v[1][0]), "=f"(v[1][1]), "=f"(v[1][2]), "=f"(v[1][3]),
        "=f"(v[2][0]), "=f"(v[2][1]), "=f"(v[2][2]), "=f"(v[2][3]), "=f"(v[3][0]), "=f"(v[3][1]), "=f"(v[3][2]), "=f"(v[3][3]),
        "=f"(v[4][0]), "=f"(v[4][1]), "=f"(v[4][2]), "=f"(v[4][3])
      : "r"(taddr), "r"(taddr + 32u), "r"(taddr + 64u), "r"(taddr + 96u), "r"(taddr + 128u)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]), "f"(v[10]),
      "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st4u(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, float a) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "f"(a) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n\ttcgen05.wait::ld.sync.aligned;"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st2u(uint32_t taddr, uint32_t a, uint32_t b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// round-to-nearest tf32 in one instruction (SASS F2FP.SATFINITE.TF32.F32; cvt.rna is an add and a mask)
__device__ __forceinline__ uint32_t cvt_tf32(float x) {
  uint32_t r;
  asm("cvt.rn.satfinite.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
// x = hi + lo for two values at once, hi and lo both rounded to nearest tf32 (the tensor core would truncate their 13 low bits, a
// one-sided error that adds up over the layers; these operands set the loss values): 2.5 instructions per value with the
// one-instruction conversion
__device__ __forceinline__ void split2(float2 x, uint32_t& h0, uint32_t& h1, uint32_t& l0, uint32_t& l1) {
  h0 = cvt_tf32(x.x);
  h1 = cvt_tf32(x.y);
  const float2 lo = fma2(make_float2(__uint_as_float(h0), __uint_as_float(h1)), bc2(-1.0f), x);
  l0 = cvt_tf32(lo.x);
  l1 = cvt_tf32(lo.y);
}
// two neighbouring neurons as bf16 pairs: p1 = (bf16(x0) | bf16(x1) << 16), p2 the same of the remainders
__device__ __forceinline__ void bf16_pair(float2 x, uint32_t& p1, uint32_t& p2) {
  p1 = pack_bf16(x.x, x.y);
  // remainder x - b1 by the mixed-precision FMA (SASS FHFMA.BF16 with a half selector: p1 is not unpacked; full FP32 rate)
  float2 r;
  asm("{\n.reg .b16 lo, hi, m1;\nmov.b32 {lo, hi}, %2;\nmov.b16 m1, 0xBF80;\n"
      "fma.rn.f32.bf16 %0, lo, m1, %3;\nfma.rn.f32.bf16 %1, hi, m1, %4;\n}\n"
      : "=f"(r.x), "=f"(r.y)
      : "r"(p1), "f"(x.x), "f"(x.y));
  p2 = pack_bf16(r.x, r.y);
}
__device__ __forceinline__ float2 bf16_unpair(uint32_t p1, uint32_t p2) {
  float2 r;   // b1 + b2 with b1 taken by the mixed-precision add's half selector (SASS FHADD.BF16)
  asm("{\n.reg .b16 lo, hi;\nmov.b32 {lo, hi}, %2;\nadd.rn.f32.bf16 %0, lo, %3;\nadd.rn.f32.bf16 %1, hi, %4;\n}\n"
      : "=f"(r.x), "=f"(r.y)
      : "r"(p1), "f"(__uint_as_float(p2 << 16)), "f"(__uint_as_float(p2 & 0xFFFF0000u)));
  return r;
}

// sums of V per-lane values over the 32 lanes of a warp by halving exchanges: V = 8 m values v[m * n8 + t] (n8 = 0..7) end
// up in the lanes with (lane & 3) == 0 as v[t], t < m, of n8 = lane >> 2   (~V shuffles instead of 5 V)
template <int V, int S>
__device__ __forceinline__ void xr_step(float* v, int lane) {
  const bool up = (lane & S) != 0;
#pragma unroll
  for (int i = 0; i < V / 2; ++i) {
    const float keep = up ? v[i + V / 2] : v[i];
    const float send = up ? v[i] : v[i + V / 2];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, S);
  }
}
template <int M>
__device__ __forceinline__ void xreduce8(float (&v)[8 * M], int lane) {
  xr_step<8 * M, 16>(v, lane);
  xr_step<4 * M, 8>(v, lane);
  xr_step<2 * M, 4>(v, lane);
#pragma unroll
  for (int i = 0; i < M; ++i) {
    v[i] += __shfl_xor_sync(0xffffffffu, v[i], 2);
    v[i] += __shfl_xor_sync(0xffffffffu, v[i], 1);
  }
}

// hi / lo images of a half-octet (4 neurons = 2 pairs) x C channels of this thread's row into tensor memory
template <int C>
__device__ __forceinline__ void emit_operand(const float2 (&v)[C][2], uint32_t tm_a) {
#pragma unroll
  for (int c = 0; c < C; ++c) {
    uint32_t h[4], l[4];
    split2(v[c][0], h[0], h[1], l[0], l[1]);
    split2(v[c][1], h[2], h[3], l[2], l[3]);
    tmem_st4u(tm_a + 64u * c, h[0], h[1], h[2], h[3]);
    tmem_st4u(tm_a + 64u * c + 32u, l[0], l[1], l[2], l[3]);
  }
}
// the adjoint GEMMs take the SAME bf16 pairs the weight-gradient images are made of as their operand (kind::f16, A in tensor
// memory: a 32-bit column holds two neighbouring neurons): part b1 of channel c at columns 64c + j / 2, part b2 at 64c + 32 + j / 2
template <int C>
__device__ __forceinline__ void emit_pairs(const uint2 (&p1)[C], const uint2 (&p2)[C], uint32_t tm_a) {
#pragma unroll
  for (int c = 0; c < C; ++c) {
    tmem_st2u(tm_a + 64u * c, p1[c].x, p1[c].y);
    tmem_st2u(tm_a + 64u * c + 32u, p2[c].x, p2[c].y);
  }
}
// bf16-pair images of the same values: rows (c, p), 4 neurons = 8 bytes per part
template <int C>
__device__ __forceinline__ void pack_images(const float2 (&v)[C][2], uint2 (&p1)[C], uint2 (&p2)[C]) {
#pragma unroll
  for (int c = 0; c < C; ++c) {
    bf16_pair(v[c][0], p1[c].x, p2[c].x);
    bf16_pair(v[c][1], p1[c].y, p2[c].y);
  }
}
// A side of a weight-gradient batch (a_2, a_1): the 16-byte atom of a row is [b1 of 4 neurons | b2 of the same 4 neurons] -- one
// conflict-free STS.128 per channel (the M index of the MMA is ours to order: row 8 hq + e of the accumulator = neuron 4 hq + e % 4,
// part e / 4, hq = half-octet).  The Z side below keeps one atom per (octet, part), which the two MMAs of a k-block select by stride.
template <int C>
__device__ __forceinline__ void store_images_a(uint8_t* img_thr_g, const uint2 (&p1)[C], const uint2 (&p2)[C]) {
#pragma unroll
  for (int c = 0; c < C; ++c) *reinterpret_cast<uint4*>(img_thr_g + c * 16384) = make_uint4(p1[c].x, p1[c].y, p2[c].x, p2[c].y);
}
// Z side (z-bar_3, z-bar_2): the two MMAs of a k-block take the b1 atoms and the b2 atoms by stride, so an atom holds one part of 8
// neurons -- the 4 + 4 neurons this thread owns in the two halves of a phase (N index 8 h + e of the accumulator = neuron
// 8 u + 4 v + e % 4 + 16 (e / 4) of warp h = 2 u + v): again one conflict-free STS.128 per channel and part.
template <int C>
__device__ __forceinline__ void store_images_z(uint8_t* img_thr_h, const uint2 (&p1)[2][C], const uint2 (&p2)[2][C]) {
#pragma unroll
  for (int c = 0; c < C; ++c) {
    *reinterpret_cast<uint4*>(img_thr_h + c * 16384) = make_uint4(p1[0][c].x, p1[0][c].y, p1[1][c].x, p1[1][c].y);
    *reinterpret_cast<uint4*>(img_thr_h + c * 16384 + 128) = make_uint4(p2[0][c].x, p2[0][c].y, p2[1][c].x, p2[1][c].y);
  }
}

template <int D, int O, bool TRAIN>
__global__ void __launch_bounds__(TcCfg<D, O>::THREADS, 1)
fused_tc_kernel(const float* __restrict__ params, const SegDev* __restrict__ segs, int n_segs, int total_tiles,
                float* __restrict__ ws, int ws_stride, int n_terms_total, int params_aligned) {
  using Cfg = TcCfg<D, O>;
  constexpr int C = Cfg::C, H = 32, P = Cfg::P, SX = Cfg::SX, SY = Cfg::SY, TW = Cfg::TERM_WORDS;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* imgX = smem + Cfg::OFF_X;
  uint8_t* imgY = smem + Cfg::OFF_Y;
  float* a1buf = reinterpret_cast<float*>(smem + Cfg::OFF_A1);
  float* wimg = reinterpret_cast<float*>(smem + Cfg::OFF_W);
  float* tot = reinterpret_cast<float*>(smem + Cfg::OFF_TOT);
  float* small = reinterpret_cast<float*>(smem + Cfg::OFF_SMALL);
  float* sK1 = small + Cfg::S_K1;
  float* sB = small + Cfg::S_B;
  float* sKo = small + Cfg::S_KO;
  float* sBo = small + Cfg::S_BO;
  float* sg_all = reinterpret_cast<float*>(smem + Cfg::OFF_SG);
  float* ssq_all = reinterpret_cast<float*>(smem + Cfg::OFF_SSQ);
  float* sterm = reinterpret_cast<float*>(smem + Cfg::OFF_TERM);
  int* segt = reinterpret_cast<int*>(smem + Cfg::OFF_SEGT);
  long long* sseg = reinterpret_cast<long long*>(smem + Cfg::OFF_SEG);    // [si][4]: (chunk_begin | n_terms << 32), n, pts, y_out
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 12);

  const int tid = threadIdx.x, nthr = Cfg::THREADS;
  const int lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform for the compiler: tensor-memory addresses stay in uniform registers

  // ---- stage the parameter vector (one TMA bulk copy + tail), build the operand images, stage the terms ---------------
  {
    TC_STAGE(0);
    float* raw = reinterpret_cast<float*>(imgX);
    const uint32_t bulk_bytes = params_aligned ? ((uint32_t)(P * 4) & ~15u) : 0u;
    if (tid == 0) {
#pragma unroll
      for (int g = 0; g < 4; ++g) mbar_init(&bar[B_AREADY + g], 8);        // the 8 warps that produce neuron octet g
      mbar_init(&bar[B_DLOADED], Cfg::NEPI);
      mbar_init(&bar[B_DFULL], 1);
      mbar_init(&bar[B_IMG], Cfg::NEPI);
      mbar_init(&bar[B_WDONE], 1);
      mbar_init(&bar[B_STAGE], 1);
      fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0 && bulk_bytes) {
      mbar_expect_tx(&bar[B_STAGE], bulk_bytes);
      tma_bulk_g2s(raw, params, bulk_bytes, &bar[B_STAGE]);
    }
    for (int i = (int)(bulk_bytes / 4) + tid; i < P; i += nthr) raw[i] = __ldg(params + i);
    // the launch's loss terms in a compact form (coefficients of the O x C output jets first): read once per tile and
    // thread from shared memory instead of through 15 dependent global loads per term
    for (int si = tid; si < n_segs; si += nthr) {
      sseg[4 * si + 0] = (long long)(unsigned)segs[si].chunk_begin | ((long long)segs[si].n_terms << 32);
      sseg[4 * si + 1] = segs[si].n;
      sseg[4 * si + 2] = (long long)reinterpret_cast<uintptr_t>(segs[si].pts);
      sseg[4 * si + 3] = (long long)reinterpret_cast<uintptr_t>(segs[si].y_out);
    }
    __syncthreads();
    TC_STAGE(1);
    if (tid == 0) {                            // first staged term of each segment: prefix sum over the table in shared memory
      int t0 = 0;
      for (int si = 0; si < n_segs; ++si) {
        segt[si] = t0;
        t0 += (int)(sseg[4 * si] >> 32);
      }
      tslot[1] = (uint32_t)t0;                 // terms of THIS launch (n_terms_total counts the whole plan)
    }
    __syncthreads();
    {
      // one flat pass over (term, word): every thread issues its global loads at once instead of one round per segment
      const int launch_terms = (int)tslot[1];
      for (int idx = tid; idx < launch_terms * TW; idx += nthr) {
        const int gt = idx / TW, k = idx % TW;
        int si = 0;
        while (si + 1 < n_segs && gt >= segt[si + 1]) ++si;
        const int t0 = segt[si], t = gt - t0;
        const TermDev* T = segs[si].terms + t;
        float val = 0.f;
        if (k < O * C) val = T->coef[k % O][k / O];          // flat index c * O + o: the order of the exchanged output jets
        else if (k == Cfg::T_CONV) val = T->conv;
        else if (k == Cfg::T_RHS_SCALE) val = T->rhs_scale;
        else if (k == Cfg::T_SCALE) val = T->scale;
        else if (k == Cfg::T_FLAGS) val = __int_as_float((T->conv_k & 0xff) | ((T->kind & 0xff) << 8) | ((T->train & 0xff) << 16));
        else if (k == Cfg::T_OUT) val = __int_as_float(T->out_index);
        else if (k == Cfg::T_RHS) val = __uint_as_float((uint32_t)(reinterpret_cast<uintptr_t>(T->rhs) & 0xffffffffu));
        else if (k == Cfg::T_RHS + 1) val = __uint_as_float((uint32_t)(reinterpret_cast<uintptr_t>(T->rhs) >> 32));
        else if (k == Cfg::T_SIGN) val = __uint_as_float((uint32_t)(reinterpret_cast<uintptr_t>(T->sign) & 0xffffffffu));
        else if (k == Cfg::T_SIGN + 1) val = __uint_as_float((uint32_t)(reinterpret_cast<uintptr_t>(T->sign) >> 32));
        sterm[(t0 + t) * TW + k] = val;
      }
    }
    TC_STAGE(2);
    if (bulk_bytes) pinn::mbar_wait(&bar[B_STAGE], 0);
    __syncthreads();
    TC_STAGE(3);
    // K-major operand images of the 32 x 32 matrices (rows n, contraction index k):
    //   forward  D[p][j] = sum_k a[p][k] K_l[k][j]:  B[n = j][k]      = K_l[k][j]   tf32 hi / lo (umma::tile_offset, SBO = 1024)
    //   adjoint  D[p][k] = sum_j z[p][j] K_l[k][j]:  B[n = k][kk = j] = K_l[k][j]   bf16 pair w1 / w2 (SBO = 512)
    for (int idx = tid; idx < 2 * H * H; idx += nthr) {
      const int l = idx / (H * H), r = (idx / H) % H, c = idx % H;   // r = row of K_l (k), c = column (j)
      const float w = raw[Cfg::offK(l + 2) + r * H + c];
      const uint32_t hi = (__float_as_uint(w) + 0x1000u) & 0xFFFFE000u;
      const uint32_t lo = __float_as_uint(w - __uint_as_float(hi)) + 0x1000u;
      const uint32_t of = umma::tile_offset(c, r, 1024) / 4;
      float* base = wimg + l * 2048;   // hi | lo
      base[of] = __uint_as_float(hi);
      base[1024 + of] = __uint_as_float(lo);
      // adjoint: bf16 pair w = w1 + w2, K-major 16-bit core matrices (8 rows x 8 elements): rows n = r, contraction index kk = c
      const uint32_t ob = (uint32_t)(r >> 3) * 512u + (uint32_t)(c >> 3) * 128u + (uint32_t)(r & 7) * 16u + (uint32_t)(c & 7) * 2u;
      const __nv_bfloat16 w1 = __float2bfloat16_rn(w);
      const __nv_bfloat16 w2 = __float2bfloat16_rn(w - __bfloat162float(w1));
      uint8_t* bb = smem + Cfg::OFF_WB + l * 4096;   // w1 | w2
      *reinterpret_cast<__nv_bfloat16*>(bb + ob) = w1;
      *reinterpret_cast<__nv_bfloat16*>(bb + 2048 + ob) = w2;
    }
    for (int idx = tid; idx < D * H; idx += nthr) sK1[idx] = raw[idx];
    for (int idx = tid; idx < 3 * H; idx += nthr) {
      const int l = idx / H, j = idx % H;
      sB[idx] = (l == 0) ? raw[D * H + j] : raw[Cfg::offK(l + 1) + H * H + j];
    }
    for (int idx = tid; idx < H * 4; idx += nthr) {     // K_out as neuron PAIRS: sKo[j / 2][o][j % 2] -- a thread's (j, j + 1) weights of
      const int j = idx >> 2, o = idx & 3;              // one output are an aligned register pair of its 16-byte loads (packed FFMA2 operand)
      sKo[(j >> 1) * 8 + o * 2 + (j & 1)] = (o < O) ? raw[Cfg::OFF_KO + j * O + o] : 0.f;
    }
    if (tid < 4) sBo[tid] = (tid < O) ? raw[Cfg::OFF_BO + tid] : 0.f;
    for (int idx = tid; idx < 2 * H * Cfg::TOT_LD; idx += nthr) tot[idx] = 0.f;
    for (int idx = tid; idx < Cfg::NEPI * Cfg::SG_FLOATS; idx += nthr) sg_all[idx] = 0.f;
    for (int idx = tid; idx < 4 * kMaxLaunchTerms; idx += nthr) ssq_all[idx] = 0.f;
    TC_STAGE(4);
    if (warp == Cfg::NEPI) umma::tmem_alloc<512>(tslot);
    umma::fence_proxy_async_smem();
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    TC_STAGE(5);
  }
  const uint32_t tmem = *tslot;
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t bar_base = smem_base + Cfg::OFF_BAR;
  auto BAR = [&](int i) -> uint32_t { return bar_base + 8u * (uint32_t)i; };

  if (warp >= Cfg::NEPI) {
    reg_dec<Cfg::AUX_REGS>();
    if (warp == Cfg::NEPI) {
      // ================================ MMA warp ===================================================================
      const bool leader = elect_one();
      constexpr uint32_t idesc_g = umma::idesc_tf32(128, 32);
      constexpr uint32_t idesc_w = idesc_bf16(64, 32, 1, 1);
      uint32_t ph_a = 0, ph_dl = 0, ph_img = 0;   // bit g of ph_a: parity of a_ready[g]
      bool first_gemm = true;
      TC_PROF_DECL
      // one forward GEMM over the 5 channel tiles: image index 0 = layer 2, 2 = layer 3 (1, 3: their adjoints, gemm_adj)
      int seq = 0;                                 // tiles this CTA has started (trace builds)
      (void)seq;
      auto gemm = [&](int image) {
        const uint64_t wh = umma::smem_desc(smem_base + Cfg::OFF_W + (image >> 1) * 8192, 128, 1024);
        const uint64_t wl = umma::smem_desc(smem_base + Cfg::OFF_W + (image >> 1) * 8192 + 4096, 128, 1024);
#pragma unroll
        for (int g = 0; g < 4; ++g) {            // k-step g contracts over neuron octet g; octets 0, 1 are ready first
          mbar_wait(BAR(B_AREADY + g), (ph_a >> g) & 1u);
          ph_a ^= 1u << g;
          if (g == 0) {
            if (!first_gemm) {                   // every epilogue warp has read the previous accumulators out of tensor memory
              mbar_wait(BAR(B_DLOADED), ph_dl);
              ph_dl ^= 1u;
            }
            first_gemm = false;
          }
          umma::fence_after_thread_sync();
          TC_PROF(1 + g);
          TC_TRACE(seq, (image == 0 ? 0 : image == 2 ? 5 : image == 3 ? 10 : 15) + g);
          if (leader) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
              const uint32_t d = tmem + Cfg::COL_D + 32u * c;
              const uint32_t ah = tmem + Cfg::COL_A + 64u * c + 8u * g, al = ah + 32u;
              mma_tf32_ts(d, al, wh + (uint64_t)(g * 16), idesc_g, g > 0 ? 1u : 0u);
              mma_tf32_ts(d, ah, wl + (uint64_t)(g * 16), idesc_g, 1u);
              mma_tf32_ts(d, ah, wh + (uint64_t)(g * 16), idesc_g, 1u);
            }
          }
          __syncwarp();
        }
        if (leader) mma_commit(BAR(B_DFULL));
        __syncwarp();
        TC_PROF(5);
        TC_TRACE(seq, (image == 0 ? 0 : image == 2 ? 5 : image == 3 ? 10 : 15) + 4);
      };
      // adjoint GEMM on the bf16 pairs (gradients only, tolerance 1e-4): z = b1 + b2, K = w1 + w2, three kind::f16 MMAs of K = 16
      // per channel and half (b2 w1 + b1 w2 + b1 w1): 30 instead of 60 MMAs, and no operand split in the epilogue
      constexpr uint32_t idesc_a = idesc_bf16(128, 32, 0, 0);
      auto gemm_adj = [&](int image) {
        const uint64_t w1 = umma::smem_desc(smem_base + Cfg::OFF_WB + (image >> 1) * 4096, 128, 512);
        const uint64_t w2 = umma::smem_desc(smem_base + Cfg::OFF_WB + (image >> 1) * 4096 + 2048, 128, 512);
#pragma unroll
        for (int k2 = 0; k2 < 2; ++k2) {         // half k2 contracts over the neuron octets 2 k2, 2 k2 + 1
#pragma unroll
          for (int g = 2 * k2; g < 2 * k2 + 2; ++g) {
            mbar_wait(BAR(B_AREADY + g), (ph_a >> g) & 1u);
            ph_a ^= 1u << g;
          }
          if (k2 == 0) {
            mbar_wait(BAR(B_DLOADED), ph_dl);
            ph_dl ^= 1u;
          }
          umma::fence_after_thread_sync();
          TC_PROF(1 + 2 * k2);
          TC_TRACE(seq, (image == 3 ? 10 : 15) + 2 * k2);
          if (leader) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
              const uint32_t d = tmem + Cfg::COL_D + 32u * c;
              const uint32_t a1 = tmem + Cfg::COL_A + 64u * c + 8u * k2, a2 = a1 + 32u;
              mma_bf16_ts(d, a2, w1 + (uint64_t)(k2 * 16), idesc_a, k2 > 0 ? 1u : 0u);
              mma_bf16_ts(d, a1, w2 + (uint64_t)(k2 * 16), idesc_a, 1u);
              mma_bf16_ts(d, a1, w1 + (uint64_t)(k2 * 16), idesc_a, 1u);
            }
          }
          __syncwarp();
        }
        if (leader) mma_commit(BAR(B_DFULL));
        __syncwarp();
        TC_PROF(5);
        TC_TRACE(seq, (image == 3 ? 10 : 15) + 4);
      };
      // weight gradient of one layer: D_w[(k-octet, part, k % 8)][j] = sum_rows a[row][k] z[row][j]
      auto wgrad = [&](int a_off, int z_off) {
        mbar_wait(BAR(B_IMG), ph_img);
        ph_img ^= 1u;
        umma::fence_after_thread_sync();
        TC_PROF(6);
        TC_TRACE(seq, a_off == Cfg::OFF_X ? 21 : 23);
        if (leader) {
          const uint64_t ad = umma::smem_desc(smem_base + a_off, 1024, 128);          // M = 64: all 8 (octet, part) blocks of a k-block
          const uint64_t z1 = umma::smem_desc(smem_base + z_off, 1024, 256);          // N = 32: the b1 blocks
          const uint64_t z2 = umma::smem_desc(smem_base + z_off + 128, 1024, 256);    //         the b2 blocks
#pragma unroll 4
          for (int ks = 0; ks < C * Cfg::TP / 16; ++ks) {
            mma_bf16_ss_keep_a(tmem + Cfg::COL_W, ad + (uint64_t)(ks * 128), z1 + (uint64_t)(ks * 128), idesc_w, ks > 0 ? 1u : 0u);
            mma_bf16_ss_reuse_a(tmem + Cfg::COL_W, ad + (uint64_t)(ks * 128), z2 + (uint64_t)(ks * 128), idesc_w, 1u);
          }
          mma_commit(BAR(B_WDONE));
        }
        __syncwarp();
        TC_PROF(7);
        TC_TRACE(seq, a_off == Cfg::OFF_X ? 20 : 22);
      };
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        gemm(0);                                  // layer 2 forward
        gemm(2);                                  // layer 3 forward
        if constexpr (TRAIN) {
          gemm_adj(3);                            // layer 3 adjoint: a-bar_2 = z-bar_3 K_3^T
          wgrad(Cfg::OFF_X, Cfg::OFF_Y);          // K-bar_3 = a_2^T z-bar_3
          gemm_adj(1);                            // layer 2 adjoint
          wgrad(Cfg::OFF_Y, Cfg::OFF_X);          // K-bar_2 = a_1^T z-bar_2
        }
        ++seq;
      }
    }
  } else {
    reg_inc<Cfg::EPI_REGS>();
    // ================================ epilogue warps ================================================================
    const int q = warp & 3, h = warp >> 2, u = h >> 1, v4 = h & 1;
    const int p = 32 * q + lane;                                   // point of the tile = tensor-memory lane
    const uint32_t tm_lane = tmem + ((uint32_t)(32 * q) << 16);
    const int img_thr_z = (p >> 3) * 1024 + (p & 7) * 16 + h * 256;    // Z-side layout: byte offset of (row (c = 0, p), atom b1 of warp h)
    const int img_thr_a = (p >> 3) * 1024 + (p & 7) * 16 + v4 * 128;   // A-side layout: atom of half-octet 2 g + v4
    float* sg = sg_all + warp * Cfg::SG_FLOATS;
    float* ssq = ssq_all + (warp & 3) * kMaxLaunchTerms;
    uint32_t ph_df = 0, ph_wd = 0;
    bool w_pending = false;                                        // a weight-gradient MMA batch not yet drained
    float gb2acc[8], gb3acc[8];                                    // bias gradients of layers 2, 3: this thread's point share, all tiles
#pragma unroll
    for (int i = 0; i < 8; ++i) gb2acc[i] = gb3acc[i] = 0.f;
    TC_PROF_DECL

    // accumulator of the finished weight-gradient batch -> FP32 totals of layer index li (0: K_2, 1: K_3)
    auto drain_w = [&](int li) {
      mbar_wait(BAR(B_WDONE), ph_wd);
      ph_wd ^= 1u;
      umma::fence_after_thread_sync();
      {
        // the four warps of a quadrant share the drain: warp h takes columns 8h .. 8h+7 (neurons j) of the quadrant's 16 rows
        float wv[8];
        tmem_ld8(tm_lane + Cfg::COL_W + 8u * (uint32_t)h, wv);
        // quadrant q holds rows 16q..16q+15 of the M = 64 accumulator (A-side atom order): lanes 8a..8a+3 the b1 part of neurons
        // 8q + 4a .. +3 (a = 0, 1), lanes 8a+4..8a+7 their b2 part
#pragma unroll
        for (int j = 0; j < 8; ++j) wv[j] += __shfl_down_sync(0xffffffffu, wv[j], 4);
        if (lane < 16 && (lane & 4) == 0) {
          // columns 8h..8h+7 (Z-side atom order): neurons 8u + 4v .. +3 and the same + 16
          float4* t4 = reinterpret_cast<float4*>(tot + (li * 32 + 8 * q + 4 * (lane >> 3) + (lane & 3)) * Cfg::TOT_LD + 8 * u + 4 * v4);
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            float4 t = t4[4 * j];
            t.x += wv[4 * j]; t.y += wv[4 * j + 1]; t.z += wv[4 * j + 2]; t.w += wv[4 * j + 3];
            t4[4 * j] = t;
          }
        }
        umma::fence_before_thread_sync();
      }
    };
    auto arrive_group = [&](int g) {           // this warp's share of neuron octet g is in tensor memory
      tmem_wait_st();
      umma::fence_before_thread_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(B_AREADY + g));
    };
    auto arrive_dloaded = [&]() {
      umma::fence_before_thread_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(B_DLOADED));
    };

    // tile state, prepared one tile ahead (the point coordinates are a global load)
    int si = 0;
    long long pg = 0, pi = 0;
    bool valid = false;
    float x[D];
    auto prepare = [&](int tile) {             // segment table in shared memory; only the coordinates are a global load
      while (si + 1 < n_segs && tile >= (int)(unsigned)sseg[4 * (si + 1)]) ++si;
      const long long n = sseg[4 * si + 1];
      pg = (long long)(tile - (int)(unsigned)sseg[4 * si]) * Cfg::TP + p;
      valid = pg < n;
      pi = valid ? pg : n - 1;
      const float* pts = reinterpret_cast<const float*>((uintptr_t)sseg[4 * si + 2]);
#pragma unroll
      for (int i = 0; i < D; ++i) x[i] = __ldg(pts + pi * D + i);
    };
    // layer 1 of the prepared tile: z = x K1 + b1, a = tanh z; jets straight into the operand of the layer-2 GEMM.
    // Runs one phase EARLY (before the layer-1 backward math of the previous tile), so that GEMM is done when its
    // epilogue starts.
    auto layer1 = [&]() {
#pragma unroll
      for (int hs = 0; hs < 2; ++hs) {
        const int g = 2 * hs + u, j0 = 8 * g + 4 * v4;
        float2 v[C][2];
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {
          const int j = j0 + 2 * pr;
          float2 z = *reinterpret_cast<const float2*>(sB + j);
          float2 zd[D], a[C];
#pragma unroll
          for (int i = 0; i < D; ++i) {
            zd[i] = *reinterpret_cast<const float2*>(sK1 + i * H + j);
            z = fma2(bc2(x[i]), zd[i], z);
          }
          const float2 a0 = tanh2(z);
          if constexpr (TRAIN) {
            a1buf[j * Cfg::TP + p] = a0.x;
            a1buf[(j + 1) * Cfg::TP + p] = a0.y;
          }
          jet2_from_a0<Cfg>(a0, zd, bc2(0.f), bc2(0.f), a);
#pragma unroll
          for (int c = 0; c < C; ++c) v[c][pr] = a[c];
        }
        emit_operand<C>(v, tm_lane + Cfg::COL_A + 8u * g + 4u * v4);
        arrive_group(g);
      }
    };
    if ((int)blockIdx.x < total_tiles) {
      prepare(blockIdx.x);
      layer1();
    }
    TC_STAGE(6);
    // sum r^2 of term h of the current segment over this thread's points (the four threads of a point hold the same residuals:
    // warp h of the quadrant keeps the sums of the terms t = h mod 4)
    float sqacc = 0.f;
    int sq_si = si;
    auto flush_sq = [&]() {                    // register -> per-quadrant slot, when the segment changes
      const int nt = (int)(sseg[4 * sq_si] >> 32);
      const float tsum = reduce_warp(sqacc);
      if (lane == 0 && h < nt) ssq[__float_as_int(sterm[(segt[sq_si] + h) * TW + Cfg::T_OUT])] += tsum;
      sqacc = 0.f;
    };

    int eseq = 0;                                // tiles this CTA has started (trace builds)
    (void)eseq;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++eseq) {
      const int si_c = si;
      const long long pg_c = pg, pi_c = pi;
      const bool valid_c = valid;
      const int term0 = segt[si];
      float xc[D];
#pragma unroll
      for (int i = 0; i < D; ++i) xc[i] = x[i];
      if (si != sq_si) {
        flush_sq();
        sq_si = si;
      }
      TC_PROF(0);

      // ---- layer 2: accumulators -> tanh jets -> operand of the layer-3 GEMM (+ images of a_2 for the reverse sweep) ---
      {
        mbar_wait(BAR(B_DFULL), ph_df);
        ph_df ^= 1u;
        umma::fence_after_thread_sync();
        TC_PROF(1);
        TC_TRACE(eseq, 0);
        if constexpr (TRAIN) {
          if (w_pending) {                     // K-bar_2 batch of the previous tile: drain before image X is overwritten
            drain_w(0);
            w_pending = false;
          }
          TC_PROF(3);
        }
#pragma unroll
        for (int hs = 0; hs < 2; ++hs) {
          const int g = 2 * hs + u, j0 = 8 * g + 4 * v4;
          float d[C][4];
          tmem_ld4x5(tm_lane + Cfg::COL_D + 8u * g + 4u * v4, d);
          if (hs == 1) arrive_dloaded();
          float2 v[C][2];
#pragma unroll
          for (int pr = 0; pr < 2; ++pr) {
            const int j = j0 + 2 * pr;
            const float2 b = *reinterpret_cast<const float2*>(sB + H + j);
            float2 zd[D], a[C];
#pragma unroll
            for (int i = 0; i < D; ++i) zd[i] = make_float2(d[1 + i][2 * pr], d[1 + i][2 * pr + 1]);
            const float2 zxx = make_float2(d[1 + D][2 * pr], d[1 + D][2 * pr + 1]);
            const float2 zyy = make_float2(d[2 + D][2 * pr], d[2 + D][2 * pr + 1]);
            const float2 z0 = add2(make_float2(d[0][2 * pr], d[0][2 * pr + 1]), b);
            jet2_from_a0<Cfg>(tanh2(z0), zd, zxx, zyy, a);
#pragma unroll
            for (int c = 0; c < C; ++c) v[c][pr] = a[c];
          }
          emit_operand<C>(v, tm_lane + Cfg::COL_A + 8u * g + 4u * v4);
          arrive_group(g);                     // the GEMM only needs the operand: the images below are written under its MMAs
          TC_TRACE(eseq, 1 + hs);
          if constexpr (TRAIN) {
            uint2 p1[C], p2[C];
            pack_images<C>(v, p1, p2);
            store_images_a<C>(imgX + img_thr_a + g * 256, p1, p2);
          }
        }
      }
      TC_PROF(2);
      TC_TRACE(eseq, 3);

      // ---- layer 3 + output layer ------------------------------------------------------------------------------------
      float a3[C][8];                          // a-jets of layer 3 of this thread's 8 neurons
      float J[C][O];
      float2 Jv[8];                            // the same output jets as pairs of the flat index c * O + o (word 15 = 0)
      {
        mbar_wait(BAR(B_DFULL), ph_df);
        ph_df ^= 1u;
        umma::fence_after_thread_sync();
        TC_PROF(4);
        TC_TRACE(eseq, 4);
        float2 Jp[C][O];
#pragma unroll
        for (int c = 0; c < C; ++c)
#pragma unroll
          for (int o = 0; o < O; ++o) Jp[c][o] = make_float2(0.f, 0.f);
#pragma unroll
        for (int hs = 0; hs < 2; ++hs) {
          const int g = 2 * hs + u, j0 = 8 * g + 4 * v4;
          float d[C][4];
          tmem_ld4x5(tm_lane + Cfg::COL_D + 8u * g + 4u * v4, d);
          if (hs == 1) arrive_dloaded();
#pragma unroll
          for (int pr = 0; pr < 2; ++pr) {
            const int j = j0 + 2 * pr;
            const float2 b = *reinterpret_cast<const float2*>(sB + 2 * H + j);
            float2 zd[D], a[C];
#pragma unroll
            for (int i = 0; i < D; ++i) zd[i] = make_float2(d[1 + i][2 * pr], d[1 + i][2 * pr + 1]);
            const float2 zxx = make_float2(d[1 + D][2 * pr], d[1 + D][2 * pr + 1]);
            const float2 zyy = make_float2(d[2 + D][2 * pr], d[2 + D][2 * pr + 1]);
            const float2 z0 = add2(make_float2(d[0][2 * pr], d[0][2 * pr + 1]), b);
            jet2_from_a0<Cfg>(tanh2(z0), zd, zxx, zyy, a);
            const float4 k0 = *reinterpret_cast<const float4*>(sKo + j * 4), k1 = *reinterpret_cast<const float4*>(sKo + j * 4 + 4);
            const float2 kv[4] = {make_float2(k0.x, k0.y), make_float2(k0.z, k0.w), make_float2(k1.x, k1.y), make_float2(k1.z, k1.w)};
#pragma unroll
            for (int c = 0; c < C; ++c) {
              a3[c][4 * hs + 2 * pr] = a[c].x;
              a3[c][4 * hs + 2 * pr + 1] = a[c].y;
#pragma unroll
              for (int o = 0; o < O; ++o) Jp[c][o] = fma2(a[c], kv[o], Jp[c][o]);
            }
          }
        }
        // the other 24 neurons live in the 3 partner warps of the quadrant (same lanes): exchange the partial output jets
        // through tensor memory -- columns 64 h + 16 .. + 31 of the operand region: idle once the forward GEMM has finished, and not
        // written by the bf16-pair operands of the reverse sweep (which take columns 64 c + 0..15 and 64 c + 32..47), so in the
        // training kernel nobody has to wait for the readers; the next writer is layer 1 of the next tile, after the last adjoint GEMM
        float mine[16], part[16];
#pragma unroll
        for (int c = 0; c < C; ++c)
#pragma unroll
          for (int o = 0; o < O; ++o) mine[c * O + o] = Jp[c][o].x + Jp[c][o].y;
#pragma unroll
        for (int i = C * O; i < 16; ++i) mine[i] = 0.f;
        tmem_st16(tm_lane + Cfg::COL_A + 64u * h + 16u, mine);
        tmem_wait_st();
        umma::fence_before_thread_sync();
        named_bar_sync(1 + q, 128);
        umma::fence_after_thread_sync();
#pragma unroll
        for (int i = 0; i < 8; ++i) Jv[i] = make_float2(2 * i < O ? sBo[(2 * i) % 4] : 0.f, 2 * i + 1 < O ? sBo[(2 * i + 1) % 4] : 0.f);
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) {       // same order in all four threads of a point: identical sums
          tmem_ld16(tm_lane + Cfg::COL_A + 64u * hh + 16u, part);
#pragma unroll
          for (int i = 0; i < 8; ++i) Jv[i] = add2(Jv[i], make_float2(part[2 * i], part[2 * i + 1]));
        }
#pragma unroll
        for (int c = 0; c < C; ++c)
#pragma unroll
          for (int o = 0; o < O; ++o) J[c][o] = ((c * O + o) & 1) ? Jv[(c * O + o) >> 1].y : Jv[(c * O + o) >> 1].x;
        if constexpr (!TRAIN) {                 // forward only: layer 1 of the next tile follows at once and overwrites the columns
          umma::fence_before_thread_sync();
          named_bar_sync(1 + q, 128);
          umma::fence_after_thread_sync();
        }
      }
      if (h == 0 && valid_c && sseg[4 * si_c + 3] != 0) {
        float* y = reinterpret_cast<float*>((uintptr_t)sseg[4 * si_c + 3]);
#pragma unroll
        for (int o = 0; o < O; ++o) y[pg_c * O + o] = J[0][o];
      }
      TC_PROF(5);
      TC_TRACE(eseq, 5);

      // ---- residuals, sums of squares, adjoint of the output jets ----------------------------------------------------------
      float2 Jbv[8];                           // adjoint of the output jets, linear part, same pairs
      float jcv[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // convective part: J-bar[0][0], [0][1], [1+SX][0], [1+SY][0], [1+SX][1], [1+SY][1]
#pragma unroll
      for (int i = 0; i < 8; ++i) Jbv[i] = make_float2(0.f, 0.f);
      const int n_terms = (int)(sseg[4 * si_c] >> 32);
      // The four threads of a point SHARE the residual work: thread h evaluates the terms t = h (mod 4) -- residual, square and the
      // adjoint seed r-bar_t -- the seeds of a group of four terms travel through 4 tensor-memory columns, and every thread then
      // accumulates the adjoint of the output jets from all of them (it needs it for its own 8 neurons).
#pragma unroll 1
      for (int tb = 0; tb < n_terms; tb += 4) {
        const int t = tb + h;
        float rb = 0.f;
        if (t < n_terms) {
          float T[TW];
          {
            const float4* T4 = reinterpret_cast<const float4*>(sterm + (term0 + t) * TW);
#pragma unroll
            for (int i = 0; i < TW / 4; ++i) {
              const float4 w = T4[i];
              T[4 * i] = w.x; T[4 * i + 1] = w.y; T[4 * i + 2] = w.z; T[4 * i + 3] = w.w;
            }
          }
          const int flags = __float_as_int(T[Cfg::T_FLAGS]);
          const int ck = flags & 0xff;
          const bool abs_mean = ((flags >> 8) & 0xff) != 0;
          if (!(TRAIN && ((flags >> 16) & 0xff) == 0)) {
            float2 ra = mul2(make_float2(T[0], T[1]), Jv[0]), rc = mul2(make_float2(T[2], T[3]), Jv[1]);
#pragma unroll
            for (int i = 2; i < 8; i += 2) {
              ra = fma2(make_float2(T[2 * i], T[2 * i + 1]), Jv[i], ra);
              rc = fma2(make_float2(T[2 * i + 2], T[2 * i + 3]), Jv[i + 1], rc);
            }
            ra = add2(ra, rc);
            float r = ra.x + ra.y;
            if constexpr (O >= 2) {
              const float ukx = ck == 0 ? J[1 + SX][0] : J[1 + SX][1];
              const float uky = ck == 0 ? J[1 + SY][0] : J[1 + SY][1];
              r = fmaf(T[Cfg::T_CONV], fmaf(J[0][0], ukx, J[0][1] * uky), r);
            }
            const float* rhs = reinterpret_cast<const float*>((uintptr_t)__float_as_uint(T[Cfg::T_RHS]) |
                                                              ((uintptr_t)__float_as_uint(T[Cfg::T_RHS + 1]) << 32));
            if (rhs != nullptr) r = fmaf(-T[Cfg::T_RHS_SCALE], __ldg(rhs + pi_c), r);
            r = valid_c ? r : 0.f;
            const float sq = abs_mean ? r : r * r;
            if (tb == 0) {
              sqacc += sq;
            } else {
              const float tsum = reduce_warp(sq);
              if (lane == 0) ssq[__float_as_int(T[Cfg::T_OUT])] += tsum;
            }
            if constexpr (TRAIN) {
              rb = T[Cfg::T_SCALE] * r;
              if (abs_mean) {
                const float* sgn = reinterpret_cast<const float*>((uintptr_t)__float_as_uint(T[Cfg::T_SIGN]) |
                                                                  ((uintptr_t)__float_as_uint(T[Cfg::T_SIGN + 1]) << 32));
                rb = valid_c ? T[Cfg::T_SCALE] * __ldg(sgn) : 0.f;
              }
            }
          }
        }
        if constexpr (TRAIN) {
          // seeds of the group: columns 160 + 256 + 16 + 4 (group parity) + h -- free during the whole reverse sweep, and two groups
          // apart from their next writer (the barrier of the group in between orders them)
          const uint32_t xcol = tm_lane + Cfg::COL_A + 256u + 16u + 4u * (uint32_t)((tb >> 2) & 1);
          tmem_st1(xcol + (uint32_t)h, rb);
          tmem_wait_st();
          umma::fence_before_thread_sync();
          named_bar_sync(1 + q, 128);
          umma::fence_after_thread_sync();
          float rbv[4];
          tmem_ld4(xcol, rbv);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (tb + k < n_terms) {
              const float4* T4 = reinterpret_cast<const float4*>(sterm + (term0 + tb + k) * TW);
              const float4 w0 = T4[0], w1 = T4[1], w2 = T4[2], w3 = T4[3], w4 = T4[4];
              const float2 rb2 = bc2(rbv[k]);
              Jbv[0] = fma2(make_float2(w0.x, w0.y), rb2, Jbv[0]);
              Jbv[1] = fma2(make_float2(w0.z, w0.w), rb2, Jbv[1]);
              Jbv[2] = fma2(make_float2(w1.x, w1.y), rb2, Jbv[2]);
              Jbv[3] = fma2(make_float2(w1.z, w1.w), rb2, Jbv[3]);
              Jbv[4] = fma2(make_float2(w2.x, w2.y), rb2, Jbv[4]);
              Jbv[5] = fma2(make_float2(w2.z, w2.w), rb2, Jbv[5]);
              Jbv[6] = fma2(make_float2(w3.x, w3.y), rb2, Jbv[6]);
              Jbv[7] = fma2(make_float2(w3.z, w3.w), rb2, Jbv[7]);      // .y: conv * rb, never read
              if constexpr (O >= 2) {
                const int ck = __float_as_int(w4.z) & 0xff;             // words 16..19: rhs_scale, scale, flags, out
                const float ukx = ck == 0 ? J[1 + SX][0] : J[1 + SX][1];
                const float uky = ck == 0 ? J[1 + SY][0] : J[1 + SY][1];
                const float m = w3.w * rbv[k];                         // word 15: conv
                jcv[0] = fmaf(m, ukx, jcv[0]);
                jcv[1] = fmaf(m, uky, jcv[1]);
                const float m0 = ck == 0 ? m : 0.f, m1 = ck == 0 ? 0.f : m;
                jcv[2] = fmaf(m0, J[0][0], jcv[2]);
                jcv[3] = fmaf(m0, J[0][1], jcv[3]);
                jcv[4] = fmaf(m1, J[0][0], jcv[4]);
                jcv[5] = fmaf(m1, J[0][1], jcv[5]);
              }
            }
          }
        }
      }
      float Jb[C][O];
#pragma unroll
      for (int c = 0; c < C; ++c)
#pragma unroll
        for (int o = 0; o < O; ++o) Jb[c][o] = ((c * O + o) & 1) ? Jbv[(c * O + o) >> 1].y : Jbv[(c * O + o) >> 1].x;
      if constexpr (O >= 2) {
        Jb[0][0] += jcv[0];
        Jb[0][1] += jcv[1];
        Jb[1 + SX][0] += jcv[2];
        Jb[1 + SY][0] += jcv[3];
        Jb[1 + SX][1] += jcv[4];
        Jb[1 + SY][1] += jcv[5];
      }
      TC_PROF(6);
      TC_TRACE(eseq, 6);

      if constexpr (TRAIN) {
        // ---- output layer backward + tanh-jet backward of layer 3: z-bar_3 -> operand of the adjoint GEMM + image Y ----
#pragma unroll
        for (int o = 0; o < O; ++o) {          // b_out gradient: output o is summed by warp h == o of the quadrant
          if (h == o) {
            const float vsum = reduce_warp(Jb[0][o]);
            if (lane == 0) sg[Cfg::SG_BO + o] += vsum;
          }
        }
        {
          float gko[24];
          uint2 zq1[2][C], zq2[2][C];            // z-bar_3 as bf16 pairs: operand of the adjoint GEMM AND image Y of the weight gradient
#pragma unroll
          for (int hs = 0; hs < 2; ++hs) {
            const int g = 2 * hs + u, j0 = 8 * g + 4 * v4;
            float2 v[C][2];
#pragma unroll
            for (int pr = 0; pr < 2; ++pr) {
              const int n8 = 4 * hs + 2 * pr, j = j0 + 2 * pr;
              const float4 k0 = *reinterpret_cast<const float4*>(sKo + j * 4), k1 = *reinterpret_cast<const float4*>(sKo + j * 4 + 4);
              const float2 kv[4] = {make_float2(k0.x, k0.y), make_float2(k0.z, k0.w), make_float2(k1.x, k1.y), make_float2(k1.z, k1.w)};
              float2 aj[C], ab[C], zb[C], zdummy[D];
#pragma unroll
              for (int i = 0; i < D; ++i) zdummy[i] = bc2(0.f);
              float2 pk[O];
#pragma unroll
              for (int o = 0; o < O; ++o) pk[o] = bc2(0.f);
#pragma unroll
              for (int c = 0; c < C; ++c) {
                aj[c] = make_float2(a3[c][n8], a3[c][n8 + 1]);
                float2 b = bc2(0.f);
#pragma unroll
                for (int o = 0; o < O; ++o) {
                  b = fma2(bc2(Jb[c][o]), kv[o], b);
                  pk[o] = fma2(aj[c], bc2(Jb[c][o]), pk[o]);
                }
                ab[c] = b;
              }
#pragma unroll
              for (int o = 0; o < 3; ++o) {
                gko[3 * n8 + o] = o < O ? pk[o < O ? o : 0].x : 0.f;
                gko[3 * (n8 + 1) + o] = o < O ? pk[o < O ? o : 0].y : 0.f;
              }
              tanh_jet2_bwd<Cfg, false>(aj, zdummy, ab, zb);
              gb3acc[n8] += zb[0].x;
              gb3acc[n8 + 1] += zb[0].y;
#pragma unroll
              for (int c = 0; c < C; ++c) v[c][pr] = zb[c];
            }
            pack_images<C>(v, zq1[hs], zq2[hs]);
            emit_pairs<C>(zq1[hs], zq2[hs], tm_lane + Cfg::COL_A + 4u * g + 2u * v4);
            arrive_group(g);
            TC_TRACE(eseq, 7 + hs);
          }
          store_images_z<C>(imgY + img_thr_z, zq1, zq2);
          umma::fence_proxy_async_smem();      // images X (a_2) and Y (z-bar_3) -> visible to the weight-gradient MMAs
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(B_IMG));
          TC_TRACE(eseq, 9);
          // K_out gradient of this warp's 8 neurons: sums over its 32 points
          xreduce8<3>(gko, lane);
          if ((lane & 3) == 0) {
#pragma unroll
            for (int o = 0; o < 3; ++o) sg[Cfg::SG_KO + (lane >> 2) * 4 + o] += gko[o];
          }
        }
        TC_PROF(7);
        TC_TRACE(eseq, 10);

        // ---- layer 2 backward: a-bar_2 (accumulators) + a_2 (image X) -> z-bar_2 ------------------------------------
        {
          uint2 q1[2][C], q2[2][C];            // a_2 of this thread's 8 neurons (its own image entries): loaded ahead of the wait
#pragma unroll
          for (int hs = 0; hs < 2; ++hs)
#pragma unroll
            for (int c = 0; c < C; ++c) {
              const uint4 qq = *reinterpret_cast<const uint4*>(imgX + img_thr_a + (2 * hs + u) * 256 + c * 16384);
              q1[hs][c] = make_uint2(qq.x, qq.y);
              q2[hs][c] = make_uint2(qq.z, qq.w);
            }
          mbar_wait(BAR(B_DFULL), ph_df);
          ph_df ^= 1u;
          umma::fence_after_thread_sync();
          TC_PROF(8);
          TC_TRACE(eseq, 11);
          uint2 zp1[2][C], zp2[2][C];          // images of z-bar_2: written once the K-bar_3 MMAs have finished reading X and Y
#pragma unroll
          for (int hs = 0; hs < 2; ++hs) {
            const int g = 2 * hs + u;
            float d[C][4];
            tmem_ld4x5(tm_lane + Cfg::COL_D + 8u * g + 4u * v4, d);
            if (hs == 1) arrive_dloaded();
            float2 v[C][2];
#pragma unroll
            for (int pr = 0; pr < 2; ++pr) {
              float2 aj[C], ab[C], zb[C], zdummy[D];
#pragma unroll
              for (int i = 0; i < D; ++i) zdummy[i] = bc2(0.f);
#pragma unroll
              for (int c = 0; c < C; ++c) {
                aj[c] = bf16_unpair(pr == 0 ? q1[hs][c].x : q1[hs][c].y, pr == 0 ? q2[hs][c].x : q2[hs][c].y);
                ab[c] = make_float2(d[c][2 * pr], d[c][2 * pr + 1]);
              }
              tanh_jet2_bwd<Cfg, false>(aj, zdummy, ab, zb);
              gb2acc[4 * hs + 2 * pr] += zb[0].x;
              gb2acc[4 * hs + 2 * pr + 1] += zb[0].y;
#pragma unroll
              for (int c = 0; c < C; ++c) v[c][pr] = zb[c];
            }
            pack_images<C>(v, zp1[hs], zp2[hs]);
            emit_pairs<C>(zp1[hs], zp2[hs], tm_lane + Cfg::COL_A + 4u * g + 2u * v4);
            arrive_group(g);
            TC_TRACE(eseq, 12 + hs);
          }
          TC_PROF(9);
          TC_TRACE(eseq, 14);
          drain_w(1);                          // K-bar_3 batch finished: its accumulator -> totals; X and Y are free
          TC_PROF(10);
          TC_TRACE(eseq, 15);
          store_images_z<C>(imgX + img_thr_z, zp1, zp2);
#pragma unroll
          for (int hs = 0; hs < 2; ++hs) {
            const int g = 2 * hs + u, j0 = 8 * g + 4 * v4;
            // a_1 jets re-materialised from tanh(z1) into image Y (left operand of the K-bar_2 MMAs)
            float2 v[C][2];
#pragma unroll
            for (int pr = 0; pr < 2; ++pr) {
              const int j = j0 + 2 * pr;
              const float2 a0 = make_float2(a1buf[j * Cfg::TP + p], a1buf[(j + 1) * Cfg::TP + p]);
              float2 zd[D], a[C];
#pragma unroll
              for (int i = 0; i < D; ++i) zd[i] = *reinterpret_cast<const float2*>(sK1 + i * H + j);
              jet2_from_a0<Cfg>(a0, zd, bc2(0.f), bc2(0.f), a);
#pragma unroll
              for (int c = 0; c < C; ++c) v[c][pr] = a[c];
            }
            uint2 p1[C], p2[C];
            pack_images<C>(v, p1, p2);
            store_images_a<C>(imgY + img_thr_a + g * 256, p1, p2);
          }
          umma::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(B_IMG));
          w_pending = true;
        }
        TC_PROF(11);
        TC_TRACE(eseq, 16);
      }

      // the next tile's segment and point coordinates (global loads), then its layer 1 as soon as the operand columns of
      // tensor memory are free again
      const bool more = tile + (int)gridDim.x < total_tiles;
      if (more) prepare(tile + gridDim.x);

      if constexpr (TRAIN) {
        // ---- layer 1 backward: a-bar_1 (accumulators) + tanh(z1) -> K1 / b1 gradients --------------------------------
        float2 a1v[4];                         // tanh(z1) of this thread's 8 neurons: read before the next tile's layer 1 overwrites it
#pragma unroll
        for (int hs = 0; hs < 2; ++hs)
#pragma unroll
          for (int pr = 0; pr < 2; ++pr) {
            const int j = 8 * (2 * hs + u) + 4 * v4 + 2 * pr;
            a1v[2 * hs + pr] = make_float2(a1buf[j * Cfg::TP + p], a1buf[(j + 1) * Cfg::TP + p]);
          }
        mbar_wait(BAR(B_DFULL), ph_df);
        ph_df ^= 1u;
        umma::fence_after_thread_sync();
        TC_PROF(12);
        TC_TRACE(eseq, 17);
        float d[2][C][4];
        tmem_ld4x5(tm_lane + Cfg::COL_D + 8u * u + 4u * v4, d[0]);
        tmem_ld4x5(tm_lane + Cfg::COL_D + 8u * (2 + u) + 4u * v4, d[1]);
        arrive_dloaded();
        if (more) layer1();                    // the adjoint GEMM has released the operand columns: next tile's layer 1 -> its GEMM runs during the math below
        TC_PROF(13);
        TC_TRACE(eseq, 18);
        float gk[8 * (D + 1)];
#pragma unroll
        for (int hs = 0; hs < 2; ++hs) {
          const int g = 2 * hs + u, j0 = 8 * g + 4 * v4;
#pragma unroll
          for (int pr = 0; pr < 2; ++pr) {
            const int j = j0 + 2 * pr, n8 = 4 * hs + 2 * pr;
            float2 aj[C], ab[C], zb[C], zd[D];
            aj[0] = a1v[2 * hs + pr];
#pragma unroll
            for (int c = 1; c < C; ++c) aj[c] = bc2(0.f);
#pragma unroll
            for (int i = 0; i < D; ++i) zd[i] = *reinterpret_cast<const float2*>(sK1 + i * H + j);
#pragma unroll
            for (int c = 0; c < C; ++c) ab[c] = make_float2(d[hs][c][2 * pr], d[hs][c][2 * pr + 1]);
            tanh_jet2_bwd<Cfg, true>(aj, zd, ab, zb);
#pragma unroll
            for (int i = 0; i < D; ++i) {
              gk[(D + 1) * n8 + i] = fmaf(xc[i], zb[0].x, zb[1 + i].x);
              gk[(D + 1) * (n8 + 1) + i] = fmaf(xc[i], zb[0].y, zb[1 + i].y);
            }
            gk[(D + 1) * n8 + D] = zb[0].x;
            gk[(D + 1) * (n8 + 1) + D] = zb[0].y;
          }
        }
        xreduce8<D + 1>(gk, lane);
        if ((lane & 3) == 0) {
#pragma unroll
          for (int i = 0; i < D; ++i) sg[Cfg::SG_K1 + i * 8 + (lane >> 2)] += gk[i];
          sg[Cfg::SG_B1 + (lane >> 2)] += gk[D];
        }
        TC_PROF(14);
        TC_TRACE(eseq, 19);
      } else {
        (void)a3;
        if (more) layer1();
      }
    }
    TC_STAGE(7);
    flush_sq();
    if constexpr (TRAIN) {
      if (w_pending) drain_w(0);
      xreduce8<1>(gb2acc, lane);
      xreduce8<1>(gb3acc, lane);
      if ((lane & 3) == 0) {
        sg[Cfg::SG_B2 + (lane >> 2)] += gb2acc[0];
        sg[Cfg::SG_B3 + (lane >> 2)] += gb3acc[0];
      }
    }

    // ---- CTA reduction over the epilogue warps: this CTA's workspace row (Keras order, then the term sums) -----------
    umma::fence_before_thread_sync();
    named_bar_sync(5, Cfg::NEPI * 32);
    float* row = ws + (size_t)blockIdx.x * ws_stride;
    constexpr int nepi = Cfg::NEPI * 32;
    if constexpr (TRAIN) {
      // neuron j = 8 (2 hs + u) + 4 v + e lives in the warps 4 (2u + v) + quadrant, slot n8 = 4 hs + e
      for (int idx = tid; idx < P; idx += nepi) {
        float s = 0.f;
        int j = -1, off = 0, stride = 1;
        if (idx < D * H) { j = idx % H; off = Cfg::SG_K1 + (idx / H) * 8; }
        else if (idx < D * H + H) { j = idx - D * H; off = Cfg::SG_B1; }
        else if (idx < Cfg::OFF_KO) {
          const int r = idx - Cfg::offK(2);
          const int l = r / (H * H + H), qq = r % (H * H + H);
          if (qq < H * H) s = tot[(l * 32 + qq / H) * Cfg::TOT_LD + qq % H];
          else { j = qq - H * H; off = l == 0 ? Cfg::SG_B2 : Cfg::SG_B3; }
        } else if (idx < Cfg::OFF_BO) {
          const int r = idx - Cfg::OFF_KO;
          j = r / O; off = Cfg::SG_KO + (r % O); stride = 4;
        } else {
#pragma unroll
          for (int w = 0; w < 4; ++w) s += sg_all[(4 * (idx - Cfg::OFF_BO) + w) * Cfg::SG_FLOATS + Cfg::SG_BO + (idx - Cfg::OFF_BO)];
        }
        if (j >= 0) {
          const int hh = 2 * ((j >> 3) & 1) + ((j >> 2) & 1), n8 = 4 * (j >> 4) + (j & 3);
#pragma unroll
          for (int w = 0; w < 4; ++w) s += sg_all[(4 * hh + w) * Cfg::SG_FLOATS + off + n8 * stride];
        }
        row[idx] = s;
      }
    }
    for (int t = tid; t < n_terms_total; t += nepi) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 4; ++w) s += ssq_all[w * kMaxLaunchTerms + t];
      row[ws_stride - n_terms_total + t] = s;
    }
  }

  TC_STAGE(8);
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == Cfg::NEPI) umma::tmem_dealloc<512>(tmem);
  TC_STAGE(9);
}

}  // namespace ftc
}  // namespace pinn
