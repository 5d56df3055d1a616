// Layered engine for wide/deep tanh MLPs (H = 64 or 128, any L >= 2), e.g. BASELINE config 5
// (3-128x8-3, 116 483 parameters).  The jets of every layer do not fit on chip for these sizes
// (24.6 KB per point for 8x128 with 6 channels), so activations live in an HBM workspace and the
// step is a sequence of per-layer kernels over a BATCH of points:
//
//   layer1_kernel            z1 = x K1 + b1, a-jets                                  -> Act[1]
//   fwd_layer_kernel  (l)    Act[l] = tanh-jet( Act[l-1] . K_l + b_l )               -> Act[l]
//   out_layer_kernel         J = Act[L] . K_out + b_out, residuals, sum r^2, J-bar,
//                            a-bar_L = J-bar . K_out^T, z-bar_L (in place over Act[L]), K_out/b_out grads
//   wgrad_kernel      (l)    K_l-bar += Act[l-1]^T . Zbar[l],  b_l-bar += sum Zbar[l][channel 0]
//   bwd_layer_kernel  (l)    a-bar_{l-1} = Zbar[l] . K_l^T, z-bar_{l-1} = tanh-jet-bwd  (in place over Act[l-1]);
//                            for l = 2 the epilogue reduces the K1 / b1 gradients instead
//
// Workspace layout ("tile-blocked", chosen so the same buffers can feed tcgen05 MMAs of M=128 =
// 2 channels x 64 points): Act[l][tile][c][p][k], tile = 64 consecutive points, k contiguous.
// This file holds the FP32 SIMT implementation of those kernels (register-tiled FFMA GEMMs).
#pragma once
#include "common.cuh"

namespace pinn {
namespace layered {

constexpr int kTile = 64;     // points per tile
constexpr int kKC = 32;       // K chunk staged in shared memory
constexpr int kNJ = 16;       // neurons per thread
constexpr int kThreads = 256; // 64 points x 4 neuron groups -> 64 columns per pass

template <int D, int ORDER>
struct Jet {
  static constexpr int C = n_channels(D, ORDER);
  static constexpr int SX = D - 2, SY = D - 1;
};

template <int D, int ORDER>
__device__ __forceinline__ void jet_fwd(float a0, const float (&zd)[D], float zxx, float zyy,
                                        float (&a)[Jet<D, ORDER>::C]) {
  a[0] = a0;
  if constexpr (ORDER >= 1) {
    const float s = fmaf(-a0, a0, 1.0f);
#pragma unroll
    for (int i = 0; i < D; ++i) a[1 + i] = s * zd[i];
    if constexpr (ORDER >= 2) {
      const float q = -2.0f * a0 * s;
      a[1 + D] = fmaf(q * zd[D - 2], zd[D - 2], s * zxx);
      a[2 + D] = fmaf(q * zd[D - 1], zd[D - 1], s * zyy);
    }
  }
}

// z-bar from (stored a-jets, a-bar); layer 1 passes its constant pre-activation jet explicitly.
template <int D, int ORDER, bool LAYER1>
__device__ __forceinline__ void jet_bwd(const float (&aj)[Jet<D, ORDER>::C], const float (&k1)[D],
                                        const float (&ab)[Jet<D, ORDER>::C], float (&zb)[Jet<D, ORDER>::C]) {
  constexpr int SX = D - 2, SY = D - 1;
  const float a0 = aj[0];
  const float s = fmaf(-a0, a0, 1.0f);
  zb[0] = s * ab[0];
  if constexpr (ORDER >= 1) {
    const float q = -2.0f * a0 * s;
    float rs;                                            // 1/s: rcp.approx + one Newton step (s in (0, 1])
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(s));
    rs = rs * fmaf(-s, rs, 2.0f);
    rs = s > 0.0f ? rs : 0.0f;
    float zd[D];
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      zd[i] = LAYER1 ? k1[i] : aj[1 + i] * rs;
      zb[1 + i] = s * ab[1 + i];
      acc = fmaf(zd[i], ab[1 + i], acc);
    }
    if constexpr (ORDER >= 2) {
      const float zx2 = zd[SX] * zd[SX], zy2 = zd[SY] * zd[SY];
      const float zxx = LAYER1 ? 0.0f : (aj[1 + D] - q * zx2) * rs;
      const float zyy = LAYER1 ? 0.0f : (aj[2 + D] - q * zy2) * rs;
      const float qp = -2.0f * s * fmaf(-3.0f * a0, a0, 1.0f);
      zb[1 + D] = s * ab[1 + D];
      zb[2 + D] = s * ab[2 + D];
      zb[1 + SX] = fmaf(2.0f * q * zd[SX], ab[1 + D], zb[1 + SX]);
      zb[1 + SY] = fmaf(2.0f * q * zd[SY], ab[2 + D], zb[1 + SY]);
      acc = fmaf(zxx, ab[1 + D], acc);
      acc = fmaf(zyy, ab[2 + D], acc);
      zb[0] = fmaf(qp, fmaf(zx2, ab[1 + D], zy2 * ab[2 + D]), zb[0]);
    }
    zb[0] = fmaf(q, acc, zb[0]);
  }
}

// ---- layer 1 -----------------------------------------------------------------------------------
// grid: tiles; block: 256 threads = 64 points x 4 neuron groups.  Points beyond n are clamped.
template <int D, int H, int ORDER>
__global__ void __launch_bounds__(kThreads) layer1_kernel(const float* __restrict__ params, const float* __restrict__ pts,
                                                          long long n, long long p_begin, float* __restrict__ act1) {
  constexpr int C = Jet<D, ORDER>::C;
  constexpr int PJ = kThreads / H;           // points handled concurrently; lanes run over neurons (coalesced)
  const int j = threadIdx.x % H, pj = threadIdx.x / H;
  const long long tile = blockIdx.x;
  const float* K1 = params;
  float zd[D];
#pragma unroll
  for (int i = 0; i < D; ++i) zd[i] = __ldg(K1 + i * H + j);
  const float b = __ldg(params + D * H + j);
  float* out = act1 + (size_t)tile * C * kTile * H;
  for (int p = pj; p < kTile; p += PJ) {
    long long gp = p_begin + tile * kTile + p;
    if (gp >= n) gp = n - 1;
    float z = b;
#pragma unroll
    for (int i = 0; i < D; ++i) z = fmaf(__ldg(pts + gp * D + i), zd[i], z);
    float a[C];
    jet_fwd<D, ORDER>(tanh_accurate(z), zd, 0.f, 0.f, a);
#pragma unroll
    for (int c = 0; c < C; ++c) out[((size_t)c * kTile + p) * H + j] = a[c];
  }
}

// ---- shared GEMM core ----------------------------------------------------------------------------
// acc[c][jj] += sum_k Atile[c][p][k] * W[k][n0 + jg*16 + jj]     (W row-major [H][H])
template <int C, int H>
__device__ __forceinline__ void tile_gemm(const float* __restrict__ Atile, const float* __restrict__ W, int n0,
                                          float (&acc)[C][kNJ], float* sA, float* sW) {
  const int tid = threadIdx.x, p = tid & 63, jg = tid >> 6;
  constexpr int SA = kKC + 1;
  for (int k0 = 0; k0 < H; k0 += kKC) {
    __syncthreads();
    for (int idx = tid; idx < C * kTile * kKC; idx += kThreads) {
      const int row = idx / kKC, kk = idx % kKC;
      sA[row * SA + kk] = Atile[(size_t)row * H + k0 + kk];
    }
    for (int idx = tid; idx < kKC * 64; idx += kThreads) {
      const int kk = idx >> 6, nn = idx & 63;
      sW[idx] = __ldg(W + (size_t)(k0 + kk) * H + n0 + nn);
    }
    __syncthreads();
#pragma unroll 4
    for (int kk = 0; kk < kKC; ++kk) {
      float a[C];
#pragma unroll
      for (int c = 0; c < C; ++c) a[c] = sA[(c * kTile + p) * SA + kk];
      float w[kNJ];
#pragma unroll
      for (int v = 0; v < kNJ / 4; ++v) {
        const float4 t = *reinterpret_cast<const float4*>(sW + kk * 64 + jg * kNJ + 4 * v);
        w[4 * v] = t.x; w[4 * v + 1] = t.y; w[4 * v + 2] = t.z; w[4 * v + 3] = t.w;
      }
#pragma unroll
      for (int c = 0; c < C; ++c)
#pragma unroll
        for (int jj = 0; jj < kNJ; ++jj) acc[c][jj] = fmaf(a[c], w[jj], acc[c][jj]);
    }
  }
}

template <int C>
constexpr int gemm_smem_bytes() { return (C * kTile * (kKC + 1) + kKC * 64 + 4 * 128) * 4; }

// ---- forward hidden layer ------------------------------------------------------------------------
template <int D, int H, int ORDER>
__global__ void __launch_bounds__(kThreads) fwd_layer_kernel(const float* __restrict__ W /* K_l [H][H] */,
                                                             const float* __restrict__ bias,
                                                             const float* __restrict__ act_in, float* __restrict__ act_out) {
  constexpr int C = Jet<D, ORDER>::C;
  extern __shared__ __align__(16) float smem[];
  float* sA = smem;
  float* sW = smem + C * kTile * (kKC + 1);
  const int p = threadIdx.x & 63, jg = threadIdx.x >> 6;
  const size_t toff = (size_t)blockIdx.x * C * kTile * H;
  for (int n0 = 0; n0 < H; n0 += 64) {
    float acc[C][kNJ];
#pragma unroll
    for (int jj = 0; jj < kNJ; ++jj) {
      acc[0][jj] = __ldg(bias + n0 + jg * kNJ + jj);
#pragma unroll
      for (int c = 1; c < C; ++c) acc[c][jj] = 0.f;
    }
    tile_gemm<C, H>(act_in + toff, W, n0, acc, sA, sW);
#pragma unroll
    for (int jj = 0; jj < kNJ; ++jj) {
      float zd[D], a[C], zxx = 0.f, zyy = 0.f;
#pragma unroll
      for (int i = 0; i < D; ++i) zd[i] = ORDER >= 1 ? acc[(ORDER >= 1) ? 1 + i : 0][jj] : 0.f;
      if constexpr (ORDER >= 2) { zxx = acc[1 + D][jj]; zyy = acc[2 + D][jj]; }
      jet_fwd<D, ORDER>(tanh_accurate(acc[0][jj]), zd, zxx, zyy, a);
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c][jj] = a[c];
    }
    float* out = act_out + toff;
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int v = 0; v < kNJ / 4; ++v)
        *reinterpret_cast<float4*>(out + ((size_t)c * kTile + p) * H + n0 + jg * kNJ + 4 * v) =
            make_float4(acc[c][4 * v], acc[c][4 * v + 1], acc[c][4 * v + 2], acc[c][4 * v + 3]);
  }
}

// ---- backward hidden layer: z-bar_{l-1} in place over Act[l-1]; for FIRST also K1/b1 gradients ------
template <int D, int H, int ORDER, bool FIRST>
__global__ void __launch_bounds__(kThreads) bwd_layer_kernel(const float* __restrict__ WT /* K_l^T [H][H] */,
                                                             const float* __restrict__ zbar_in, float* __restrict__ act_prev,
                                                             const float* __restrict__ params, const float* __restrict__ pts,
                                                             long long n, long long p_begin, int n_tiles,
                                                             float* __restrict__ grad) {
  constexpr int C = Jet<D, ORDER>::C;
  extern __shared__ __align__(16) float smem[];
  float* sA = smem;
  float* sW = smem + C * kTile * (kKC + 1);
  float* sG = sW + kKC * 64;                 // FIRST: [(1 + D)][H] block-level K1 / b1 gradient partials
  const int p = threadIdx.x & 63, jg = threadIdx.x >> 6;
  if constexpr (FIRST) {
    for (int i = threadIdx.x; i < (1 + D) * H; i += kThreads) sG[i] = 0.f;
  }
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const size_t toff = (size_t)tile * C * kTile * H;
    float x[D];
    if constexpr (FIRST) {
      long long gp = p_begin + (long long)tile * kTile + p;
      if (gp >= n) gp = n - 1;
#pragma unroll
      for (int i = 0; i < D; ++i) x[i] = __ldg(pts + gp * D + i);
    }
    for (int n0 = 0; n0 < H; n0 += 64) {
      float acc[C][kNJ];
#pragma unroll
      for (int c = 0; c < C; ++c)
#pragma unroll
        for (int jj = 0; jj < kNJ; ++jj) acc[c][jj] = 0.f;
      tile_gemm<C, H>(zbar_in + toff, WT, n0, acc, sA, sW);
      float* ap = act_prev + toff;
#pragma unroll
      for (int jj = 0; jj < kNJ; ++jj) {
        const int j = n0 + jg * kNJ + jj;
        float aj[C], ab[C], zb[C], k1[D];
#pragma unroll
        for (int c = 0; c < C; ++c) {
          aj[c] = (FIRST && c > 0) ? 0.f : ap[((size_t)c * kTile + p) * H + j];
          ab[c] = acc[c][jj];
        }
#pragma unroll
        for (int i = 0; i < D; ++i) k1[i] = FIRST ? __ldg(params + i * H + j) : 0.f;
        jet_bwd<D, ORDER, FIRST>(aj, k1, ab, zb);
        if constexpr (!FIRST) {
#pragma unroll
          for (int c = 0; c < C; ++c) ap[((size_t)c * kTile + p) * H + j] = zb[c];
        } else {
          // reduce over the 32 points of the warp, then one shared-memory atomic per (warp, parameter)
          float vb = zb[0];
          float vk[D];
#pragma unroll
          for (int i = 0; i < D; ++i) vk[i] = x[i] * zb[0] + (ORDER >= 1 ? zb[(ORDER >= 1) ? 1 + i : 0] : 0.f);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            vb += __shfl_xor_sync(0xffffffffu, vb, o);
#pragma unroll
            for (int i = 0; i < D; ++i) vk[i] += __shfl_xor_sync(0xffffffffu, vk[i], o);
          }
          if ((threadIdx.x & 31) == 0) {
            atomicAdd(sG + D * H + j, vb);
#pragma unroll
            for (int i = 0; i < D; ++i) atomicAdd(sG + i * H + j, vk[i]);
          }
        }
      }
    }
  }
  if constexpr (FIRST) {
    __syncthreads();
    for (int i = threadIdx.x; i < (1 + D) * H; i += kThreads) atomicAdd(grad + i, sG[i]);   // [K1 | b1] are contiguous
  }
}

// ---- weight gradient of a hidden layer: split over tiles, register tile 8x8 per thread -------------
// gK[i][j] += sum_{rows} A[row][i] * Z[row][j]; gb[j] += sum_{p} Z[c=0][p][j]
template <int C, int H>
__global__ void __launch_bounds__(kThreads) wgrad_kernel(const float* __restrict__ act_prev, const float* __restrict__ zbar,
                                                         int n_tiles, float* __restrict__ gK, float* __restrict__ gb) {
  // block computes a 128x128 (or HxH) output with 256 threads: thread (ti, tj) owns rows i = ti*8.., cols j = tj*8..
  constexpr int TT = H / 16;   // H=128 -> 8x8 per thread; H=64 -> 4x4
  extern __shared__ __align__(16) float smem[];
  float* sA = smem;            // [32 rows][H]
  float* sZ = smem + 32 * H;   // [32 rows][H]
  const int tid = threadIdx.x, ti = tid >> 4, tj = tid & 15;
  float acc[TT][TT];
#pragma unroll
  for (int a = 0; a < TT; ++a)
#pragma unroll
    for (int b = 0; b < TT; ++b) acc[a][b] = 0.f;
  float bsum = 0.f;    // threads tid < H accumulate the bias gradient of column tid
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const float* A = act_prev + (size_t)tile * C * kTile * H;
    const float* Z = zbar + (size_t)tile * C * kTile * H;
    for (int r0 = 0; r0 < C * kTile; r0 += 32) {
      __syncthreads();
      for (int idx = tid; idx < 32 * H / 4; idx += kThreads) {
        reinterpret_cast<float4*>(sA)[idx] = reinterpret_cast<const float4*>(A + (size_t)r0 * H)[idx];
        reinterpret_cast<float4*>(sZ)[idx] = reinterpret_cast<const float4*>(Z + (size_t)r0 * H)[idx];
      }
      __syncthreads();
      if (r0 < kTile && tid < H) {
#pragma unroll 8
        for (int r = 0; r < 32; ++r) bsum += sZ[r * H + tid];
      }
#pragma unroll 4
      for (int r = 0; r < 32; ++r) {
        float av[TT], zv[TT];
#pragma unroll
        for (int a = 0; a < TT; ++a) av[a] = sA[r * H + ti + 16 * a];
#pragma unroll
        for (int b = 0; b < TT; ++b) zv[b] = sZ[r * H + tj + 16 * b];
#pragma unroll
        for (int a = 0; a < TT; ++a)
#pragma unroll
          for (int b = 0; b < TT; ++b) acc[a][b] = fmaf(av[a], zv[b], acc[a][b]);
      }
    }
  }
#pragma unroll
  for (int a = 0; a < TT; ++a)
#pragma unroll
    for (int b = 0; b < TT; ++b) atomicAdd(gK + (size_t)(ti + 16 * a) * H + tj + 16 * b, acc[a][b]);
  if (tid < H) atomicAdd(gb + tid, bsum);
}

// ---- output layer + residuals + adjoint, one warp per point ---------------------------------------
template <int D, int H, int O, int ORDER, bool TRAIN>
__global__ void __launch_bounds__(256) out_layer_kernel(const float* __restrict__ params, int off_ko, const SegDev* __restrict__ seg_ptr,
                                                        long long p_begin, int n_tiles, float* __restrict__ actL,
                                                        float* __restrict__ grad, float* __restrict__ sumsq) {
  constexpr int C = Jet<D, ORDER>::C;
  constexpr int KL = H / 32;         // k's per lane
  constexpr int SX = D - 2, SY = D - 1;
  const SegDev* __restrict__ seg = seg_ptr;
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const long long n = seg->n;
  const float* Ko = params + off_ko;
  const float* bo = Ko + H * O;
  float ko[KL][O];
#pragma unroll
  for (int q = 0; q < KL; ++q)
#pragma unroll
    for (int o = 0; o < O; ++o) ko[q][o] = __ldg(Ko + (lane + 32 * q) * O + o);
  float gko[KL][O];
#pragma unroll
  for (int q = 0; q < KL; ++q)
#pragma unroll
    for (int o = 0; o < O; ++o) gko[q][o] = 0.f;
  float gbo[O];
#pragma unroll
  for (int o = 0; o < O; ++o) gbo[o] = 0.f;
  float sq[kMaxTerms];
#pragma unroll
  for (int t = 0; t < kMaxTerms; ++t) sq[t] = 0.f;
  const long long total_pts = (long long)n_tiles * kTile;
  for (long long lp = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); lp < total_pts;
       lp += (long long)gridDim.x * warps_per_block) {
    const long long gp = p_begin + lp;
    const bool valid = gp < n;
    const long long tile = lp / kTile;
    const int p = (int)(lp % kTile);
    float* base = actL + (size_t)tile * C * kTile * H;
    float a[C][KL];
    float J[C][O];
#pragma unroll
    for (int c = 0; c < C; ++c) {
#pragma unroll
      for (int o = 0; o < O; ++o) J[c][o] = 0.f;
#pragma unroll
      for (int q = 0; q < KL; ++q) {
        a[c][q] = base[((size_t)c * kTile + p) * H + lane + 32 * q];
#pragma unroll
        for (int o = 0; o < O; ++o) J[c][o] = fmaf(a[c][q], ko[q][o], J[c][o]);
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int o = 0; o < O; ++o) {
        float v = J[c][o];
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
        J[c][o] = v + (c == 0 ? __ldg(bo + o) : 0.f);
      }
    if (seg->y_out != nullptr && valid && lane == 0) {
#pragma unroll
      for (int o = 0; o < O; ++o) seg->y_out[gp * O + o] = J[0][o];
    }
    float Jb[C][O];
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int o = 0; o < O; ++o) Jb[c][o] = 0.f;
    const int n_terms = seg->n_terms;
#pragma unroll
    for (int t = 0; t < kMaxTerms; ++t) {
      if (t >= n_terms) break;
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
      sq[t] = fmaf(r, r, sq[t]);
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
      for (int o = 0; o < O; ++o) gbo[o] += Jb[0][o];
#pragma unroll
      for (int q = 0; q < KL; ++q) {
        float aj[C], ab[C], zb[C], k1[D];
#pragma unroll
        for (int i = 0; i < D; ++i) k1[i] = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          aj[c] = a[c][q];
          float b = 0.f;
#pragma unroll
          for (int o = 0; o < O; ++o) {
            b = fmaf(Jb[c][o], ko[q][o], b);
            gko[q][o] = fmaf(aj[c], Jb[c][o], gko[q][o]);
          }
          ab[c] = b;
        }
        jet_bwd<D, ORDER, false>(aj, k1, ab, zb);
#pragma unroll
        for (int c = 0; c < C; ++c) base[((size_t)c * kTile + p) * H + lane + 32 * q] = zb[c];
      }
    }
  }
  // flush: every lane holds distinct K_out rows; b_out / sum r^2 are identical across lanes (lane 0 adds)
  if constexpr (TRAIN) {
#pragma unroll
    for (int q = 0; q < KL; ++q)
#pragma unroll
      for (int o = 0; o < O; ++o) atomicAdd(grad + off_ko + (lane + 32 * q) * O + o, gko[q][o]);
    if (lane == 0) {
#pragma unroll
      for (int o = 0; o < O; ++o) atomicAdd(grad + off_ko + H * O + o, gbo[o]);
    }
  }
  if (lane == 0) {
    const int n_terms = seg->n_terms;
#pragma unroll
    for (int t = 0; t < kMaxTerms; ++t)
      if (t < n_terms && (!TRAIN || seg->terms[t].train)) atomicAdd(sumsq + seg->terms[t].out_index, sq[t]);
  }
}

// K^T copies for the backward GEMMs: WT[l][j][i] = K_l[i][j]
__global__ void transpose_weights_kernel(const float* __restrict__ params, int H, int n_layers, int off0, int stride,
                                         float* __restrict__ wt) {
  const int l = blockIdx.y;
  const float* K = params + off0 + (size_t)l * stride;
  float* T = wt + (size_t)l * H * H;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < H * H; idx += gridDim.x * blockDim.x) {
    const int i = idx / H, j = idx % H;
    T[(size_t)j * H + i] = K[idx];
  }
}

}  // namespace layered
}  // namespace pinn
