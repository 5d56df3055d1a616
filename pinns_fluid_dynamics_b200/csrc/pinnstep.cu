// libpinnstep.so -- C ABI (include/pinnstep.h) over the sm_100a kernels.
#include "../../include/pinnstep.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "fused_fp32.cuh"
#include "fused_tc.cuh"
#include "p2p.cuh"
#include "bfgs.cuh"
#include "layered_fp32.cuh"
#include "layered_tc.cuh"

using namespace pinn;

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(expr)                                                                         \
  do {                                                                                         \
    cudaError_t e_ = (expr);                                                                   \
    if (e_ != cudaSuccess)                                                                     \
      return fail(PINN_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

// The opt-in shared-memory size of a kernel is a per-DEVICE attribute: one bit per device ordinal records where it has been set
// (a process may drive several GPUs).
struct DeviceOnce {
  std::atomic<unsigned long long> done{0};
  bool needed(int device) const { return device < 0 || device >= 64 || !((done.load(std::memory_order_acquire) >> device) & 1ull); }
  void mark(int device) { if (device >= 0 && device < 64) done.fetch_or(1ull << device, std::memory_order_release); }
};

// Launches go to the thread's current device: the entry points that run a plan switch to the plan's device for the call and
// restore the caller's afterwards (no-op in the usual one-GPU-per-process case).
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int device) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != device) {
      err = cudaSetDevice(device);
      switched = err == cudaSuccess;
    }
  }
  ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
struct LaunchTable {
  int order = 0;
  int n_segs = 0;
  int total_chunks = 0;
  SegDev* segs_dev = nullptr;
  std::vector<SegDev> segs_host;
  // layered tensor-core engine: the tiles of every segment of this order, set after set
  TileDev* tiles_dev = nullptr;
  std::vector<TileDev> tiles_host;
};

struct pinn_plan {
  pinn_mlp_desc mlp{};
  int device = 0;
  int num_sms = 0;
  int64_t P = 0;
  int T = 0;
  std::vector<pinn_pointset_desc> sets;
  std::vector<int> term_base;            // first output slot of each set
  LaunchTable train[3], eval[3];         // by derivative order
  LaunchTable prepass[3];                // sets that carry a |mean| training term (forward pre-pass for its sign)
  bool has_abs_mean = false;
  int fused_tensor = 0;                  // fused engines: 1 = hidden-layer GEMMs on the tensor path (fixed at plan creation)
  bool tcgen05 = false;                  // fused_tcgen05: the order-2 launch of the 2-32x3-3 network runs fused_tc_kernel (128-point tiles)
  float* signs = nullptr;                // [T] +-1 per term slot (only |mean| slots are read)
  float* prepass_out = nullptr;          // [P + T] scratch of the pre-pass
  float* ws = nullptr;                   // [rows_max][P + T]
  int rows_max = 0;
  size_t ws_bytes = 0;
  int last_launches = 0;
  const char* engine = "fused_fp32";
  bool layered = false;          // wide/deep networks: per-layer kernels over an HBM workspace
  float* act = nullptr;          // layered: [L][C_max][batch][H] jets
  float* wt = nullptr;           // layered: K_l^T copies, l = 2..L
  long long batch = 0;           // layered: points per batch (multiple of 64)
  bool tc = false;               // layered_tf32x3: tcgen05 hidden layers (H = 128)
  float* wimg = nullptr;         // tc: hi/lo operand images of K_l and K_l^T
  size_t layer_stride = 0;       // tc: floats between per-layer jet buffers
  float* tc_rows = nullptr;      // tc: one [P + T] row per CTA of the accumulating kernels (summed in a fixed order: no atomics)
  int tc_row_count = 0;
  size_t tc_row_stride = 0;
  bool timing = false;
  cudaEvent_t ev0[3] = {nullptr, nullptr, nullptr}, ev1[3] = {nullptr, nullptr, nullptr};
  bool ev_valid[3] = {false, false, false};
};

typedef void (*fused_fn)(const float*, const SegDev*, int, int, float*, int, int, int);

struct FusedKernel {
  fused_fn fn;
  int smem_bytes;
  int nw;
};

template <int D, int H, int L, int O, int ORDER, bool TRAIN, bool TENSOR>
static FusedKernel make_kernel() {
  using Cfg = FusedCfg<D, H, L, O, ORDER, TENSOR>;
  return FusedKernel{(fused_fn)fused_step_kernel<D, H, L, O, ORDER, TRAIN, TENSOR>, Cfg::SMEM_BYTES, Cfg::NW};
}

template <int D, int H, int L, int O, bool TENSOR>
static bool pick_order(int order, bool train, FusedKernel* k) {
  switch (order) {
    case 0: *k = train ? make_kernel<D, H, L, O, 0, true, TENSOR>() : make_kernel<D, H, L, O, 0, false, TENSOR>(); return true;
    case 1: *k = train ? make_kernel<D, H, L, O, 1, true, TENSOR>() : make_kernel<D, H, L, O, 1, false, TENSOR>(); return true;
    case 2: *k = train ? make_kernel<D, H, L, O, 2, true, TENSOR>() : make_kernel<D, H, L, O, 2, false, TENSOR>(); return true;
  }
  return false;
}

// H = 32 runs its hidden-layer GEMMs on the tensor path unless PINN_ENGINE=fused_fp32 asks for the FP32 FFMA2 build of
// the same kernel (cross-check of the tensor path, and the tighter-accuracy option: DESIGN.md section 6)
static bool fused_tensor_path(const pinn_mlp_desc& m) {
  const char* force = getenv("PINN_ENGINE");
  return m.width == 32 && pinn::FusedCfg<2, 32, 3, 3, 2, true>::MMA && !(force && strcmp(force, "fused_fp32") == 0);
}

static bool pick_kernel(const pinn_mlp_desc& m, int order, bool train, FusedKernel* k, int tensor = -1) {
  if (m.n_hidden != 3) return false;
  const bool tp = tensor < 0 ? fused_tensor_path(m) : tensor != 0;
  if (m.in_dim == 2 && m.width == 32 && m.out_dim == 3)
    return tp ? pick_order<2, 32, 3, 3, true>(order, train, k) : pick_order<2, 32, 3, 3, false>(order, train, k);
  if (m.in_dim == 3 && m.width == 32 && m.out_dim == 3)
    return tp ? pick_order<3, 32, 3, 3, true>(order, train, k) : pick_order<3, 32, 3, 3, false>(order, train, k);
  if (m.in_dim == 2 && m.width == 20 && m.out_dim == 1) return pick_order<2, 20, 3, 1, false>(order, train, k);
  if (m.in_dim == 2 && m.width == 20 && m.out_dim == 3) return pick_order<2, 20, 3, 3, false>(order, train, k);   // pressmean variant
  return false;
}

// fused_tcgen05 serves the second-order launch of the 2-32x3-3 network (C = 5 channels: the tensor-memory budget of fused_tc.cuh);
// every other launch of such a plan stays on the warp-level tensor path
static bool tcgen05_supported(const pinn_mlp_desc& m) {
  const char* force = getenv("PINN_ENGINE");
  if (force && (strcmp(force, "fused_fp32") == 0 || strcmp(force, "fused_tf32x3") == 0)) return false;
  return m.in_dim == 2 && m.width == 32 && m.n_hidden == 3 && m.out_dim == 3;
}
static int table_chunk_points(const pinn_plan* p, int order) { return (p->tcgen05 && order == 2) ? ftc::TcCfg<2, 3>::TP : kChunk; }

static int64_t param_count(const pinn_mlp_desc& m) {
  const int64_t d = m.in_dim, H = m.width, L = m.n_hidden, O = m.out_dim;
  return d * H + H + (L - 1) * (H * H + H) + H * O + O;
}

// mode 0: every term (pinn_loss); 1: training terms; 2: only the sets with a |mean| training term (sign pre-pass)
static int table_chunk_points(const pinn_plan* p, int order);

static int build_table(pinn_plan* p, int order, int mode, LaunchTable* out) {
  const int chunk_pts = table_chunk_points(p, order);
  const bool train_only = mode != 0;
  std::vector<SegDev> segs;
  int chunk = 0;
  for (size_t s = 0; s < p->sets.size(); ++s) {
    const pinn_pointset_desc& ps = p->sets[s];
    if (ps.deriv_order != order || ps.n_local <= 0) continue;
    if (mode == 2) {
      bool any = false;
      for (int t = 0; t < ps.n_terms; ++t) any = any || (ps.terms[t].train && ps.terms[t].kind == PINN_TERM_ABS_MEAN);
      if (!any) continue;
    }
    SegDev sd;
    memset(&sd, 0, sizeof(sd));
    int nt = 0;
    for (int t = 0; t < ps.n_terms; ++t) {
      const pinn_term_desc& td = ps.terms[t];
      if (train_only && !td.train) continue;
      if (mode == 2 && td.kind != PINN_TERM_ABS_MEAN) continue;
      TermDev& d = sd.terms[nt++];
      memcpy(d.coef, td.coef, sizeof(d.coef));
      d.conv = td.conv;
      d.conv_k = td.conv_k;
      d.rhs_scale = td.rhs_scale;
      d.rhs = td.rhs_dev;
      d.train = td.train;
      d.out_index = p->term_base[s] + t;
      d.kind = td.kind;
      d.sign = p->signs ? p->signs + d.out_index : nullptr;
      const double denom = td.normalization * (double)td.n_global;
      // d/dtheta of w/(nu N) sum r^2 is (2w/(nu N)) sum r dr; of w/(nu N) |sum r| it is (sign w/(nu N)) sum dr
      d.scale = (td.train && denom != 0.0) ? (float)((td.kind == PINN_TERM_ABS_MEAN ? 1.0 : 2.0) * td.weight / denom) : 0.f;
    }
    if (nt == 0) continue;
    sd.pts = ps.points_dev;
    sd.y_out = nullptr;
    sd.n = ps.n_local;
    sd.n_terms = nt;
    sd.chunk_begin = chunk;
    sd.n_chunks = (int)((ps.n_local + chunk_pts - 1) / chunk_pts);
    chunk += sd.n_chunks;
    segs.push_back(sd);
  }
  out->order = order;
  out->segs_host = segs;
  out->n_segs = (int)segs.size();
  out->total_chunks = chunk;
  out->segs_dev = nullptr;
  if (!segs.empty()) {
    CUDA_TRY(cudaMalloc(&out->segs_dev, segs.size() * sizeof(SegDev)));
    CUDA_TRY(cudaMemcpy(out->segs_dev, segs.data(), segs.size() * sizeof(SegDev), cudaMemcpyHostToDevice));
  }
  return PINN_OK;
}


// ------------------------------------------------------------------------------------------------
// layered engine: host orchestration
// ------------------------------------------------------------------------------------------------
static bool layered_supported(const pinn_mlp_desc& m) {
  return (m.in_dim == 2 || m.in_dim == 3) && (m.width == 64 || m.width == 128) && m.n_hidden >= 2 && m.out_dim == 3;
}

static int layered_alloc(pinn_plan* p) {
  const pinn_mlp_desc& m = p->mlp;
  long long max_n = 0;
  for (const auto& ps : p->sets) max_n = ps.n_local > max_n ? ps.n_local : max_n;
  const long long per_point = (long long)m.n_hidden * kMaxCh * m.width * 4;   // bytes of jets per point
  long long batch = (4LL << 30) / per_point;                                  // <= 4 GiB of activations
  batch = (batch / 64) * 64;
  const long long need = ((max_n + 63) / 64) * 64;
  if (batch > need) batch = need;
  if (batch < 64) batch = 64;
  p->batch = batch;
  p->ws_bytes = (size_t)batch * per_point + (size_t)(m.n_hidden - 1) * m.width * m.width * 4;
  if (cudaMalloc(&p->act, (size_t)batch * per_point) != cudaSuccess)
    return fail(PINN_E_ALLOC, "cannot allocate %lld activation bytes", batch * per_point);
  if (cudaMalloc(&p->wt, (size_t)(m.n_hidden - 1) * m.width * m.width * 4) != cudaSuccess)
    return fail(PINN_E_ALLOC, "cannot allocate transposed weights");
  return PINN_OK;
}

template <int D, int H, int O, int ORDER>
static int layered_run_set(pinn_plan* p, const float* params, float* out, cudaStream_t st, bool train,
                           const SegDev& seg, const SegDev* seg_dev, int* launches) {
  using namespace pinn::layered;
  constexpr int C = n_channels(D, ORDER);
  const int L = p->mlp.n_hidden;
  const size_t layer_stride = (size_t)p->batch * kMaxCh * H;   // floats between Act[l] buffers
  const int smem = gemm_smem_bytes<C>();
  static DeviceOnce attrs;
  if (attrs.needed(p->device)) {
    CUDA_TRY(cudaFuncSetAttribute((const void*)fwd_layer_kernel<D, H, ORDER>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUDA_TRY(cudaFuncSetAttribute((const void*)bwd_layer_kernel<D, H, ORDER, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUDA_TRY(cudaFuncSetAttribute((const void*)bwd_layer_kernel<D, H, ORDER, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attrs.mark(p->device);
  }
  const int off_ko = D * H + H + (L - 1) * (H * H + H);
  for (long long b0 = 0; b0 < seg.n; b0 += p->batch) {
    const long long nb = (seg.n - b0 < p->batch) ? seg.n - b0 : p->batch;
    const int tiles = (int)((nb + kTile - 1) / kTile);
    layer1_kernel<D, H, ORDER><<<tiles, kThreads, 0, st>>>(params, seg.pts, seg.n, b0, p->act);
    for (int l = 2; l <= L; ++l) {
      const float* K = params + D * H + H + (size_t)(l - 2) * (H * H + H);
      fwd_layer_kernel<D, H, ORDER><<<tiles, kThreads, smem, st>>>(K, K + H * H, p->act + (l - 2) * layer_stride,
                                                                  p->act + (l - 1) * layer_stride);
    }
    int grid_out = (tiles * kTile + 7) / 8;
    if (grid_out > 8 * p->num_sms) grid_out = 8 * p->num_sms;
    float* actL = p->act + (size_t)(L - 1) * layer_stride;
    if (train)
      out_layer_kernel<D, H, O, ORDER, true><<<grid_out, 256, 0, st>>>(params, off_ko, seg_dev, b0, tiles, actL, out, out + p->P);
    else
      out_layer_kernel<D, H, O, ORDER, false><<<grid_out, 256, 0, st>>>(params, off_ko, seg_dev, b0, tiles, actL, out, out + p->P);
    *launches += L + 1;
    if (train) {
      for (int l = L; l >= 2; --l) {
        float* gK = out + D * H + H + (size_t)(l - 2) * (H * H + H);
        int gw = tiles < 2 * p->num_sms ? tiles : 2 * p->num_sms;
        wgrad_kernel<C, H><<<gw, kThreads, 2 * 32 * H * 4, st>>>(p->act + (l - 2) * layer_stride, p->act + (l - 1) * layer_stride,
                                                                 tiles, gK, gK + H * H);
        const float* WT = p->wt + (size_t)(l - 2) * H * H;
        if (l > 2) {
          bwd_layer_kernel<D, H, ORDER, false><<<tiles, kThreads, smem, st>>>(WT, p->act + (l - 1) * layer_stride,
                                                                             p->act + (l - 2) * layer_stride, params, seg.pts,
                                                                             seg.n, b0, tiles, out);
        } else {
          int gb = tiles < 2 * p->num_sms ? tiles : 2 * p->num_sms;
          bwd_layer_kernel<D, H, ORDER, true><<<gb, kThreads, smem, st>>>(WT, p->act + (l - 1) * layer_stride,
                                                                         p->act + (l - 2) * layer_stride, params, seg.pts,
                                                                         seg.n, b0, tiles, out);
        }
        *launches += 2;
      }
    }
    CUDA_TRY(cudaGetLastError());
  }
  return PINN_OK;
}

template <int D, int H, int O>
static int layered_run_t(pinn_plan* p, const float* params, float* out, cudaStream_t st, bool train, int* launches) {
  for (int o = 2; o >= 0; --o) {
    const LaunchTable& lt = train ? p->train[o] : p->eval[o];
    for (int s = 0; s < lt.n_segs; ++s) {
      int rc;
      if (o == 2) rc = layered_run_set<D, H, O, 2>(p, params, out, st, train, lt.segs_host[s], lt.segs_dev + s, launches);
      else if (o == 1) rc = layered_run_set<D, H, O, 1>(p, params, out, st, train, lt.segs_host[s], lt.segs_dev + s, launches);
      else rc = layered_run_set<D, H, O, 0>(p, params, out, st, train, lt.segs_host[s], lt.segs_dev + s, launches);
      if (rc != PINN_OK) return rc;
    }
  }
  return PINN_OK;
}

static int run_layered(pinn_plan* p, const float* params, float* out, cudaStream_t st, bool train) {
  const pinn_mlp_desc& m = p->mlp;
  const int i_begin = train ? 0 : (int)p->P;
  CUDA_TRY(cudaMemsetAsync(out + i_begin, 0, sizeof(float) * (size_t)(p->P + p->T - i_begin), st));
  int launches = 0;
  if (train) {
    dim3 g(32, m.n_hidden - 1);
    pinn::layered::transpose_weights_kernel<<<g, 256, 0, st>>>(params, m.width, m.n_hidden - 1, m.in_dim * m.width + m.width,
                                                                m.width * m.width + m.width, p->wt);
    ++launches;
  }
  int rc = PINN_E_INVALID;
  if (m.in_dim == 3 && m.width == 128) rc = layered_run_t<3, 128, 3>(p, params, out, st, train, &launches);
  else if (m.in_dim == 2 && m.width == 128) rc = layered_run_t<2, 128, 3>(p, params, out, st, train, &launches);
  else if (m.in_dim == 3 && m.width == 64) rc = layered_run_t<3, 64, 3>(p, params, out, st, train, &launches);
  else if (m.in_dim == 2 && m.width == 64) rc = layered_run_t<2, 64, 3>(p, params, out, st, train, &launches);
  p->last_launches = launches;
  return rc;
}

// ------------------------------------------------------------------------------------------------
// tensor-core layered engine (H = 128): host orchestration
// ------------------------------------------------------------------------------------------------
static bool tc_supported(const pinn_mlp_desc& m) {
  return (m.in_dim == 2 || m.in_dim == 3) && m.width == 128 && m.n_hidden >= 2 && m.out_dim == 3;
}

static int tc_alloc(pinn_plan* p) {
  const pinn_mlp_desc& m = p->mlp;
  long long max_n = 0;
  for (const auto& ps : p->sets) max_n = ps.n_local > max_n ? ps.n_local : max_n;
  const long long per_point = (long long)m.n_hidden * kMaxCh * tc::kH * 4;   // bytes of jets per point
  long long budget = 16LL << 30;   // 16 GiB of the 180 GB: 4 -> 16 GiB measured 185 -> 180 ms per 4 M points (fewer, longer launches)
  if (const char* e = getenv("PINN_TC_WORKSPACE_MB")) {
    const long long mb = atoll(e);
    if (mb > 0) budget = mb << 20;
  }
  constexpr long long kAlign = 1680;   // lcm of the tile sizes (40, 48, 56, 80, 240 points)
  long long batch = (budget / per_point / kAlign) * kAlign;
  const long long need = ((max_n + kAlign - 1) / kAlign) * kAlign;
  if (batch > need) batch = need;
  if (batch < kAlign) batch = kAlign;
  p->batch = batch;
  p->layer_stride = (size_t)(6 * batch + 256) * tc::kH;
  const size_t act_bytes = (size_t)m.n_hidden * p->layer_stride * 4;
  const size_t img_bytes = (size_t)(m.n_hidden - 1) * tc::kLayerImgFloats * 4;
  p->ws_bytes = act_bytes + img_bytes;
  if (cudaMalloc(&p->act, act_bytes) != cudaSuccess) return fail(PINN_E_ALLOC, "cannot allocate %zu activation bytes", act_bytes);
  if (cudaMalloc(&p->wimg, img_bytes) != cudaSuccess) return fail(PINN_E_ALLOC, "cannot allocate weight images");
  // per-CTA accumulation rows: the widest accumulating launch (tc_out_layer) has 4 CTAs per SM
  p->tc_row_count = 4 * p->num_sms;
  p->tc_row_stride = (size_t)((p->P + p->T + 31) / 32) * 32;
  const size_t row_bytes = (size_t)p->tc_row_count * p->tc_row_stride * 4;
  if (cudaMalloc(&p->tc_rows, row_bytes) != cudaSuccess) return fail(PINN_E_ALLOC, "cannot allocate %zu bytes of accumulation rows", row_bytes);
  p->ws_bytes += row_bytes;
  return PINN_OK;
}

template <int D, int ORDER>
static int tc_set_attrs() {
  using S = tc::LayerSmem<tc::Geo<D, ORDER>::NR>;
  CUDA_TRY(cudaFuncSetAttribute((const void*)tc::tc_layer<D, ORDER, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
  CUDA_TRY(cudaFuncSetAttribute((const void*)tc::tc_layer<D, ORDER, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
  CUDA_TRY(cudaFuncSetAttribute((const void*)tc::tc_layer<D, ORDER, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
  CUDA_TRY(cudaFuncSetAttribute((const void*)tc::tc_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kWgSmem));
  return PINN_OK;
}

static constexpr int tc_points_per_tile(int d, int order) {
  const int C = n_channels(d, order);
  return C == 6 ? 40 : C == 5 ? 48 : C == 4 ? 56 : C == 3 ? 80 : 240;   // tc::Geo<D, ORDER>::P
}
static_assert(tc_points_per_tile(3, 2) == tc::Geo<3, 2>::P && tc_points_per_tile(3, 1) == tc::Geo<3, 1>::P &&
              tc_points_per_tile(3, 0) == tc::Geo<3, 0>::P && tc_points_per_tile(2, 2) == tc::Geo<2, 2>::P &&
              tc_points_per_tile(2, 1) == tc::Geo<2, 1>::P && tc_points_per_tile(2, 0) == tc::Geo<2, 0>::P,
              "host tile table and kernel tile geometry disagree");

// (re)build the tile table of one launch table: tiles never straddle point sets
static int tc_build_tiles(LaunchTable* lt, int d) {
  if (lt->tiles_dev) { cudaFree(lt->tiles_dev); lt->tiles_dev = nullptr; }
  lt->tiles_host.clear();
  const int P = tc_points_per_tile(d, lt->order);
  for (int s = 0; s < lt->n_segs; ++s)
    for (long long b = 0; b < lt->segs_host[s].n; b += P) {
      TileDev t;
      t.seg = s; t.pad_ = 0; t.p_begin = b;
      lt->tiles_host.push_back(t);
    }
  if (!lt->tiles_host.empty()) {
    CUDA_TRY(cudaMalloc(&lt->tiles_dev, lt->tiles_host.size() * sizeof(TileDev)));
    CUDA_TRY(cudaMemcpy(lt->tiles_dev, lt->tiles_host.data(), lt->tiles_host.size() * sizeof(TileDev), cudaMemcpyHostToDevice));
  }
  return PINN_OK;
}

// one pipeline run over ALL point sets of a derivative order (batches of tiles that fit the workspace)
template <int D, int O, int ORDER>
static int tc_run_order(pinn_plan* p, const float* params, float* out, cudaStream_t st, bool train, const LaunchTable& lt,
                        int* launches) {
  using G = tc::Geo<D, ORDER>;
  using S = tc::LayerSmem<G::NR>;
  constexpr int H = tc::kH;
  const int L = p->mlp.n_hidden;
  static DeviceOnce attrs;
  if (attrs.needed(p->device)) {
    int rc = tc_set_attrs<D, ORDER>();
    if (rc != PINN_OK) return rc;
    attrs.mark(p->device);
  }
  const int off_ko = D * H + H + (L - 1) * (H * H + H);
  const long long total_tiles = (long long)lt.tiles_host.size();
  const long long batch_tiles = p->batch / G::P;
  for (long long t0 = 0; t0 < total_tiles; t0 += batch_tiles) {
    const int tiles = (int)(total_tiles - t0 < batch_tiles ? total_tiles - t0 : batch_tiles);
    const TileDev* td = lt.tiles_dev + t0;
    const int grid = tiles < p->num_sms ? tiles : p->num_sms;
    auto act = [&](int l) { return p->act + (size_t)(l - 1) * p->layer_stride; };   // jets of layer l (1-based)
    tc::tc_layer1<D, ORDER><<<tiles, 256, 0, st>>>(params, lt.segs_dev, td, act(1));
    for (int l = 2; l <= L; ++l) {
      const float* bias = params + D * H + H + (size_t)(l - 2) * (H * H + H) + H * H;
      const float* img = p->wimg + (size_t)(l - 2) * tc::kLayerImgFloats;
      tc::tc_layer<D, ORDER, 0><<<grid, tc::kLayerThreads, S::TOTAL, st>>>(img, bias, act(l - 1), act(l), params, tiles);
    }
    const int grid_out = tiles < 4 * p->num_sms ? tiles : 4 * p->num_sms;
    if (train)
      tc::tc_out_layer<D, O, ORDER, true><<<grid_out, 256, 0, st>>>(params, off_ko, lt.segs_dev, td, tiles, act(L), p->tc_rows,
                                                                    p->tc_row_stride, (int)p->P);
    else
      tc::tc_out_layer<D, O, ORDER, false><<<grid_out, 256, 0, st>>>(params, off_ko, lt.segs_dev, td, tiles, act(L), p->tc_rows,
                                                                     p->tc_row_stride, (int)p->P);
    *launches += L + 1;
    if (train) {
      const long long n_slabs = (long long)tiles * (G::NR / tc::kWgRows);
      const int gw = (int)(n_slabs < p->num_sms ? n_slabs : p->num_sms);
      for (int l = L; l >= 2; --l) {
        const size_t off_gk = (size_t)(D * H + H) + (size_t)(l - 2) * (H * H + H);
        tc::tc_wgrad<<<gw, tc::kWgThreads, tc::kWgSmem, st>>>(act(l - 1), act(l), n_slabs, G::NR, G::P, p->tc_rows, p->tc_row_stride,
                                                              off_gk);
        const float* img = p->wimg + (size_t)(l - 2) * tc::kLayerImgFloats + 2 * tc::kImgFloats;
        if (l > 2)
          tc::tc_layer<D, ORDER, 1><<<grid, tc::kLayerThreads, S::TOTAL, st>>>(img, nullptr, act(l), act(l - 1), params, tiles);
        else
          tc::tc_layer<D, ORDER, 2><<<grid, tc::kLayerThreads, S::TOTAL, st>>>(img, nullptr, act(l), act(l - 1), params, tiles);
        *launches += 2;
      }
      const int g1 = tiles < 2 * p->num_sms ? tiles : 2 * p->num_sms;
      tc::tc_layer1_grad<D, ORDER><<<g1, 256, 0, st>>>(act(1), lt.segs_dev, td, tiles, p->tc_rows, p->tc_row_stride);
      ++*launches;
    }
    CUDA_TRY(cudaGetLastError());
  }
  return PINN_OK;
}

template <int D, int O>
static int tc_run_t(pinn_plan* p, const float* params, float* out, cudaStream_t st, bool train, int* launches) {
  for (int o = 2; o >= 0; --o) {
    const LaunchTable& lt = train ? p->train[o] : p->eval[o];
    if (lt.n_segs == 0) continue;
    if (p->timing) CUDA_TRY(cudaEventRecord(p->ev0[o], st));
    int rc;
    if (o == 2) rc = tc_run_order<D, O, 2>(p, params, out, st, train, lt, launches);
    else if (o == 1) rc = tc_run_order<D, O, 1>(p, params, out, st, train, lt, launches);
    else rc = tc_run_order<D, O, 0>(p, params, out, st, train, lt, launches);
    if (rc != PINN_OK) return rc;
    if (p->timing) {
      CUDA_TRY(cudaEventRecord(p->ev1[o], st));
      p->ev_valid[o] = true;
    }
  }
  return PINN_OK;
}

static int run_tc(pinn_plan* p, const float* params, float* out, cudaStream_t st, bool train) {
  const pinn_mlp_desc& m = p->mlp;
  const int i_begin = train ? 0 : (int)p->P;
  // every accumulating kernel adds into its CTAs' own rows; rows_used = the widest grid any launch of this step can have
  long long max_tiles = 1;
  for (int o = 0; o < 3; ++o) {
    const long long t = (long long)(train ? p->train[o] : p->eval[o]).tiles_host.size();
    if (t > max_tiles) max_tiles = t;
  }
  // (tc_wgrad strides over 16-row slabs: up to 15 per tile)
  const int rows_used = (int)(15 * max_tiles < p->tc_row_count ? 15 * max_tiles : p->tc_row_count);
  CUDA_TRY(cudaMemsetAsync(p->tc_rows, 0, sizeof(float) * (size_t)rows_used * p->tc_row_stride, st));
  int launches = 0;
  dim3 g(16, m.n_hidden - 1);
  tc::tc_prep_weights<<<g, 256, 0, st>>>(params, m.in_dim * m.width + m.width, m.width * m.width + m.width, p->wimg);
  ++launches;
  int rc = PINN_E_INVALID;
  if (m.in_dim == 3) rc = tc_run_t<3, 3>(p, params, out, st, train, &launches);
  else if (m.in_dim == 2) rc = tc_run_t<2, 3>(p, params, out, st, train, &launches);
  if (rc == PINN_OK) {
    // fixed-order sum of the rows -> out[i_begin .. P + T): the result does not depend on scheduling (bit-reproducible)
    const int n_out = (int)(p->P + p->T);
    if (n_out > i_begin)      // (a values-only forward through a temporary plan has no terms: nothing to sum)
      finalize_rows_range_kernel<<<(n_out - i_begin + 31) / 32, 1024, 0, st>>>(p->tc_rows, rows_used, (int)p->tc_row_stride, i_begin, n_out, out);
    CUDA_TRY(cudaGetLastError());
    ++launches;
  }
  p->last_launches = launches;
  return rc;
}

#ifdef PINN_TC_PROFILE
extern "C" int pinn_debug_tc_prof(unsigned long long* out32, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out32, tc::g_tc_prof, sizeof(unsigned long long) * 32);
  if (reset) { unsigned long long z[32] = {0}; cudaMemcpyToSymbol(tc::g_tc_prof, z, sizeof(z)); }
  return 0;
}
#endif
extern "C" int pinn_version(void) { return PINN_VERSION; }
extern "C" const char* pinn_last_error(void) { return g_err; }

extern "C" int pinn_plan_create(const pinn_mlp_desc* mlp, const pinn_pointset_desc* sets, int32_t n_sets,
                                int32_t device, pinn_plan** out) {
  if (!mlp || !out || n_sets < 0 || (n_sets > 0 && !sets)) return fail(PINN_E_INVALID, "null argument");
  *out = nullptr;
  if (mlp->in_dim < 2 || mlp->in_dim > PINN_MAX_DIM || mlp->out_dim < 1 || mlp->out_dim > PINN_MAX_OUT ||
      mlp->width < 4 || mlp->n_hidden < 1)
    return fail(PINN_E_INVALID, "unsupported MLP shape d=%d H=%d L=%d O=%d", mlp->in_dim, mlp->width, mlp->n_hidden,
                mlp->out_dim);
  FusedKernel probe;
  const bool use_fused = pick_kernel(*mlp, 0, true, &probe);
  const char* force = getenv("PINN_ENGINE");
  const bool want_fp32 = force && strcmp(force, "layered_fp32") == 0;
  const bool use_tc = !use_fused && tc_supported(*mlp) && !want_fp32;
  const bool use_layered = !use_fused && !use_tc && layered_supported(*mlp);
  if (!use_fused && !use_layered && !use_tc)
    return fail(PINN_E_INVALID,
                "no engine for MLP d=%d H=%d L=%d O=%d (fused_fp32: 2-20x3-1, 2-20x3-3, 2-32x3-3, 3-32x3-3; layered_tf32x3: "
                "d in {2,3}, H = 128, L >= 2, O = 3; layered_fp32: d in {2,3}, H in {64,128}, L >= 2, O = 3)", mlp->in_dim, mlp->width, mlp->n_hidden, mlp->out_dim);
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(PINN_E_ARCH, "device %d is sm_%d%d; this library is sm_100a only", device, prop.major, prop.minor);
  pinn_plan* p = new (std::nothrow) pinn_plan();
  if (!p) return fail(PINN_E_ALLOC, "out of host memory");
  p->mlp = *mlp;
  p->device = device;
  p->num_sms = prop.multiProcessorCount;
  p->P = param_count(*mlp);
  int T = 0;
  for (int s = 0; s < n_sets; ++s) {
    const pinn_pointset_desc& ps = sets[s];
    if (ps.n_terms < 0 || ps.n_terms > PINN_MAX_TERMS_PER_SET || ps.deriv_order < 0 || ps.deriv_order > 2 ||
        ps.n_local < 0 || (ps.n_local > 0 && !ps.points_dev)) {
      delete p;
      return fail(PINN_E_INVALID, "bad point set %d", s);
    }
    p->sets.push_back(ps);
    p->term_base.push_back(T);
    T += ps.n_terms;
  }
  if (T > kMaxLaunchTerms) {
    delete p;
    return fail(PINN_E_INVALID, "%d loss terms exceed the limit of %d", T, kMaxLaunchTerms);
  }
  p->T = T;
  if (use_fused) {
    // Small lower-order sets (boundary edges, fit points: 0.4 % of the chunks of the BASELINE configs) ride in the launch of
    // the highest derivative order instead of getting kernels of their own: their residual coefficients on the extra
    // channels are zero, the value channel is computed by the same instruction sequence, and the step loses a launch
    // and its tail.  Only when they are at most 1/16 of that launch.
    // Also when they fit into the idle part of that launch's last wave (10 k collocation points are 79 tiles of 128 points on
    // 148 SMs: the boundary tiles run beside them for free and the step is one launch instead of three).
    const bool tc_tiles = tcgen05_supported(*mlp);
    const long long per_unit = tc_tiles ? 8 : 1;                       // chunks per work unit: 128-point tile | 16-point chunk
    const long long wave = tc_tiles ? p->num_sms : 8LL * p->num_sms;   // work units in flight: one tile per CTA | one chunk per warp
    long long chunks[3] = {0, 0, 0}, units[3] = {0, 0, 0};
    for (const pinn_pointset_desc& ps : p->sets)
      if (ps.n_local > 0 && ps.n_terms > 0) {
        const long long c = (ps.n_local + kChunk - 1) / kChunk;
        chunks[ps.deriv_order] += c;
        units[ps.deriv_order] += (c + per_unit - 1) / per_unit;
      }
    int top = chunks[2] ? 2 : (chunks[1] ? 1 : 0);
    long long lower = 0, lower_units = 0;
    for (int o = 0; o < top; ++o) { lower += chunks[o]; lower_units += units[o]; }
    const bool rides_free = (units[top] + lower_units + wave - 1) / wave == (units[top] + wave - 1) / wave;
    if (top > 0 && lower > 0 && (lower * 16 <= chunks[top] || rides_free) && !(getenv("PINN_NO_PROMOTE") && atoi(getenv("PINN_NO_PROMOTE"))))
      for (pinn_pointset_desc& ps : p->sets) ps.deriv_order = top;
    p->tcgen05 = tcgen05_supported(*mlp);
  }
  for (int s = 0; s < n_sets; ++s)
    for (int t = 0; t < sets[s].n_terms; ++t) {
      const int kind = sets[s].terms[t].kind;
      if (kind != PINN_TERM_MEAN_SQUARES && kind != PINN_TERM_ABS_MEAN) {
        delete p;
        return fail(PINN_E_INVALID, "set %d term %d: unknown term kind %d", s, t, kind);
      }
      if (kind == PINN_TERM_ABS_MEAN) {
        if (!use_fused) {
          delete p;
          return fail(PINN_E_INVALID, "|mean| terms (PINN_TERM_ABS_MEAN) are served by the fused engines only");
        }
        p->has_abs_mean = p->has_abs_mean || sets[s].terms[t].train;
      }
    }
  if (p->has_abs_mean) {
    std::vector<float> ones((size_t)T, 1.f);
    if (cudaMalloc(&p->signs, sizeof(float) * (size_t)T) != cudaSuccess ||
        cudaMalloc(&p->prepass_out, sizeof(float) * (size_t)(p->P + T)) != cudaSuccess) {
      pinn_plan_destroy(p);
      return fail(PINN_E_ALLOC, "sign buffers");
    }
    CUDA_TRY(cudaMemcpy(p->signs, ones.data(), sizeof(float) * (size_t)T, cudaMemcpyHostToDevice));
  }
  for (int o = 0; o < 3; ++o) {
    int rc = build_table(p, o, 1, &p->train[o]);
    if (rc == PINN_OK) rc = build_table(p, o, 0, &p->eval[o]);
    if (rc == PINN_OK && p->has_abs_mean) rc = build_table(p, o, 2, &p->prepass[o]);
    if (rc != PINN_OK) {
      pinn_plan_destroy(p);
      return rc;
    }
  }
  if (use_tc) {
    p->tc = true;
    p->engine = "layered_tf32x3";
    int rc = tc_alloc(p);
    for (int o = 0; o < 3 && rc == PINN_OK; ++o) {
      rc = tc_build_tiles(&p->train[o], mlp->in_dim);
      if (rc == PINN_OK) rc = tc_build_tiles(&p->eval[o], mlp->in_dim);
    }
    if (rc != PINN_OK) {
      pinn_plan_destroy(p);
      return rc;
    }
    *out = p;
    return PINN_OK;
  }
  if (use_layered) {
    p->layered = true;
    p->engine = "layered_fp32";
    int rc = layered_alloc(p);
    if (rc != PINN_OK) {
      pinn_plan_destroy(p);
      return rc;
    }
    *out = p;
    return PINN_OK;
  }
  // H = 32: hidden-layer GEMMs on the warp-level tensor path (3xTF32), see fused_fp32.cuh
  p->fused_tensor = fused_tensor_path(*mlp) ? 1 : 0;
  if (p->fused_tensor) p->engine = "fused_tf32x3";
  if (p->tcgen05) {
    p->engine = "fused_tcgen05";
    using TCfg = ftc::TcCfg<2, 3>;
    cudaError_t e = cudaFuncSetAttribute((const void*)ftc::fused_tc_kernel<2, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCfg::SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute((const void*)ftc::fused_tc_kernel<2, 3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      pinn_plan_destroy(p);
      return fail(PINN_E_CUDA, "cudaFuncSetAttribute(fused_tc_kernel, smem=%d): %s", TCfg::SMEM_BYTES, cudaGetErrorString(e));
    }
  }
  p->rows_max = 3 * p->num_sms;
  p->ws_bytes = (size_t)p->rows_max * (size_t)(p->P + p->T) * sizeof(float);
  if (cudaMalloc(&p->ws, p->ws_bytes) != cudaSuccess) {
    pinn_plan_destroy(p);
    return fail(PINN_E_ALLOC, "cannot allocate %zu workspace bytes", p->ws_bytes);
  }
  // opt in to the large dynamic shared memory once, for every kernel the plan can launch
  for (int o = 0; o < 3; ++o)
    for (int tr = 0; tr < 2; ++tr) {
      FusedKernel k;
      pick_kernel(p->mlp, o, tr != 0, &k, p->fused_tensor);
      cudaError_t e = cudaFuncSetAttribute((const void*)k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, k.smem_bytes);
      if (e != cudaSuccess) {
        pinn_plan_destroy(p);
        return fail(PINN_E_CUDA, "cudaFuncSetAttribute(smem=%d): %s", k.smem_bytes, cudaGetErrorString(e));
      }
    }
  *out = p;
  return PINN_OK;
}

extern "C" int pinn_plan_destroy(pinn_plan* p) {
  if (!p) return PINN_OK;
  cudaSetDevice(p->device);
  for (int o = 0; o < 3; ++o) {
    if (p->train[o].segs_dev) cudaFree(p->train[o].segs_dev);
    if (p->eval[o].segs_dev) cudaFree(p->eval[o].segs_dev);
    if (p->train[o].tiles_dev) cudaFree(p->train[o].tiles_dev);
    if (p->eval[o].tiles_dev) cudaFree(p->eval[o].tiles_dev);
    if (p->prepass[o].segs_dev) cudaFree(p->prepass[o].segs_dev);
  }
  if (p->signs) cudaFree(p->signs);
  if (p->prepass_out) cudaFree(p->prepass_out);
  if (p->ws) cudaFree(p->ws);
  if (p->act) cudaFree(p->act);
  if (p->wt) cudaFree(p->wt);
  if (p->wimg) cudaFree(p->wimg);
  if (p->tc_rows) cudaFree(p->tc_rows);
  for (int o = 0; o < 3; ++o) {
    if (p->ev0[o]) cudaEventDestroy(p->ev0[o]);
    if (p->ev1[o]) cudaEventDestroy(p->ev1[o]);
  }
  delete p;
  return PINN_OK;
}

extern "C" int64_t pinn_plan_param_count(const pinn_plan* p) { return p ? p->P : 0; }
extern "C" int32_t pinn_plan_term_count(const pinn_plan* p) { return p ? p->T : 0; }
extern "C" size_t pinn_plan_workspace_bytes(const pinn_plan* p) { return p ? p->ws_bytes : 0; }
extern "C" const char* pinn_plan_engine(const pinn_plan* p) { return p ? p->engine : ""; }
extern "C" int32_t pinn_plan_last_launch_count(const pinn_plan* p) { return p ? p->last_launches : 0; }

extern "C" int pinn_plan_enable_timing(pinn_plan* p, int32_t on) {
  if (!p) return fail(PINN_E_INVALID, "null plan");
  if (on && !p->ev0[0]) {
    CUDA_TRY(cudaSetDevice(p->device));
    for (int o = 0; o < 3; ++o) {
      CUDA_TRY(cudaEventCreate(&p->ev0[o]));
      CUDA_TRY(cudaEventCreate(&p->ev1[o]));
    }
  }
  p->timing = on != 0;
  return PINN_OK;
}

extern "C" int pinn_plan_kernel_time_ms(pinn_plan* p, int32_t order, float* ms_out) {
  if (!p || !ms_out || order < 0 || order > 2) return fail(PINN_E_INVALID, "bad argument");
  if (!p->ev_valid[order]) return fail(PINN_E_STATE, "no timed launch recorded for derivative order %d", order);
  CUDA_TRY(cudaEventSynchronize(p->ev1[order]));
  CUDA_TRY(cudaEventElapsedTime(ms_out, p->ev0[order], p->ev1[order]));
  return PINN_OK;
}

extern "C" int pinn_plan_set_rhs(pinn_plan* p, int32_t set_index, int32_t term_index, const float* rhs_dev) {
  if (!p || set_index < 0 || set_index >= (int)p->sets.size()) return fail(PINN_E_INVALID, "bad set index");
  pinn_pointset_desc& ps = p->sets[set_index];
  if (term_index < 0 || term_index >= ps.n_terms) return fail(PINN_E_INVALID, "bad term index");
  ps.terms[term_index].rhs_dev = rhs_dev;
  // rebuild the two tables of this derivative order (rare operation; synchronous)
  CUDA_TRY(cudaSetDevice(p->device));
  const int o = ps.deriv_order;
  if (p->train[o].segs_dev) { cudaFree(p->train[o].segs_dev); p->train[o].segs_dev = nullptr; }
  if (p->eval[o].segs_dev) { cudaFree(p->eval[o].segs_dev); p->eval[o].segs_dev = nullptr; }
  if (p->prepass[o].segs_dev) { cudaFree(p->prepass[o].segs_dev); p->prepass[o].segs_dev = nullptr; }
  int rc = build_table(p, o, 1, &p->train[o]);
  if (rc == PINN_OK) rc = build_table(p, o, 0, &p->eval[o]);
  if (rc == PINN_OK && p->has_abs_mean) rc = build_table(p, o, 2, &p->prepass[o]);
  if (rc == PINN_OK && p->tc) {
    rc = tc_build_tiles(&p->train[o], p->mlp.in_dim);
    if (rc == PINN_OK) rc = tc_build_tiles(&p->eval[o], p->mlp.in_dim);
  }
  return rc;
}

// sign[t] = +-1 from the pre-pass sums (slots of other terms are never read)
__global__ void sign_kernel(const float* __restrict__ sums, float* __restrict__ sign, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) sign[i] = sums[i] < 0.f ? -1.f : 1.f;
}

// One pass of the fused engine over `tables` (train: gradient + training sums into out[0, P+T); otherwise the term
// sums only, into out[P, P+T)).
static int fused_pass(pinn_plan* p, const LaunchTable* tables, const float* params, float* out, cudaStream_t st, bool train,
                      bool timed, int* launches_out) {
  const int stride = (int)(p->P + p->T);
  int rows = 0, launches = 0;
  const int aligned = ((uintptr_t)params & 15u) == 0;
  // big derivative order first: the collocation kernel dominates
  for (int o = 2; o >= 0; --o) {
    const LaunchTable& lt = tables[o];
    if (lt.n_segs == 0) continue;
    FusedKernel k;
    if (p->tcgen05 && o == 2) {
      using TCfg = ftc::TcCfg<2, 3>;
      k = FusedKernel{train ? (fused_fn)ftc::fused_tc_kernel<2, 3, true> : (fused_fn)ftc::fused_tc_kernel<2, 3, false>, TCfg::SMEM_BYTES,
                      TCfg::THREADS / 32};
    } else if (!pick_kernel(p->mlp, o, train, &k, p->fused_tensor)) return fail(PINN_E_INVALID, "no kernel");
    int grid = lt.total_chunks < p->num_sms ? lt.total_chunks : p->num_sms;    // chunks (tiles) are dealt to CTAs first, then to warps
    if (rows + grid > p->rows_max) return fail(PINN_E_STATE, "workspace rows exhausted");
    if (timed) CUDA_TRY(cudaEventRecord(p->ev0[o], st));
    k.fn<<<grid, k.nw * 32, k.smem_bytes, st>>>(params, lt.segs_dev, lt.n_segs, lt.total_chunks,
                                                p->ws + (size_t)rows * stride, stride, p->T, aligned);
    CUDA_TRY(cudaGetLastError());
    if (timed) {
      CUDA_TRY(cudaEventRecord(p->ev1[o], st));
      p->ev_valid[o] = true;
    }
    rows += grid;
    ++launches;
  }
  const int i_begin = train ? 0 : (int)p->P;
  const int count = stride - i_begin;
  if (rows == 0) {
    CUDA_TRY(cudaMemsetAsync(out + i_begin, 0, sizeof(float) * count, st));
  } else if (count > 0) {
    finalize_rows_kernel<<<(count + 31) / 32, 1024, 0, st>>>(p->ws, rows, stride, i_begin, out);
    CUDA_TRY(cudaGetLastError());
    ++launches;
  }
  *launches_out += launches;
  return PINN_OK;
}

static int run(pinn_plan* p, const float* params, float* out, cudaStream_t st, bool train) {
  if (!p || !params || !out) return fail(PINN_E_INVALID, "null argument");
  DeviceGuard on_device(p->device);
  if (on_device.err != cudaSuccess) return fail(PINN_E_CUDA, "cannot switch to device %d: %s", p->device, cudaGetErrorString(on_device.err));
  if (p->tc) return run_tc(p, params, out, st, train);
  if (p->layered) return run_layered(p, params, out, st, train);
  int launches = 0;
  if (train && p->has_abs_mean) {
    // forward pre-pass over the sets with a |mean| term: sum r -> sign (colliding_flow_pressmean.py:176-179)
    int rc = fused_pass(p, p->prepass, params, p->prepass_out, st, false, false, &launches);
    if (rc != PINN_OK) return rc;
    sign_kernel<<<(p->T + 127) / 128, 128, 0, st>>>(p->prepass_out + p->P, p->signs, p->T);
    CUDA_TRY(cudaGetLastError());
    ++launches;
  }
  int rc = fused_pass(p, train ? p->train : p->eval, params, out, st, train, p->timing, &launches);
  if (rc != PINN_OK) return rc;
  p->last_launches = launches;
  return PINN_OK;
}

extern "C" int pinn_loss_and_grad(pinn_plan* p, const float* params_dev, float* out_dev, void* stream) {
  return run(p, params_dev, out_dev, (cudaStream_t)stream, true);
}

extern "C" int pinn_loss(pinn_plan* p, const float* params_dev, float* out_dev, void* stream) {
  return run(p, params_dev, out_dev, (cudaStream_t)stream, false);
}

extern "C" int pinn_forward(const pinn_mlp_desc* mlp, const float* params_dev, const float* points_dev, int64_t n,
                            float* y_dev, int32_t device, void* stream) {
  if (!mlp || !params_dev || (n > 0 && (!points_dev || !y_dev))) return fail(PINN_E_INVALID, "null argument");
  if (n <= 0) return PINN_OK;
  FusedKernel k;
  CUDA_TRY(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)stream;
  if (!pick_kernel(*mlp, 0, false, &k)) {
    if (!layered_supported(*mlp))
      return fail(PINN_E_INVALID, "no engine for MLP d=%d H=%d L=%d O=%d", mlp->in_dim, mlp->width, mlp->n_hidden,
                  mlp->out_dim);
    // wide/deep network: run the layered forward kernels (tensor-core engine for H = 128) over a temporary
    // workspace (synchronous free)
    const char* force = getenv("PINN_ENGINE");
    const bool use_tc = tc_supported(*mlp) && !(force && strcmp(force, "layered_fp32") == 0);
    pinn_plan tmp;
    tmp.mlp = *mlp;
    tmp.device = device;
    tmp.P = param_count(*mlp);
    CUDA_TRY(cudaDeviceGetAttribute(&tmp.num_sms, cudaDevAttrMultiProcessorCount, device));
    pinn_pointset_desc fake;
    memset(&fake, 0, sizeof(fake));
    fake.n_local = n;
    tmp.sets.push_back(fake);
    int rc = use_tc ? tc_alloc(&tmp) : layered_alloc(&tmp);
    SegDev* sd_dev = nullptr;
    if (rc == PINN_OK && cudaMalloc(&sd_dev, sizeof(SegDev)) != cudaSuccess) rc = fail(PINN_E_ALLOC, "segment descriptor");
    if (rc == PINN_OK) {
      SegDev sd;
      memset(&sd, 0, sizeof(sd));
      sd.pts = points_dev;
      sd.y_out = y_dev;
      sd.n = n;
      cudaMemcpyAsync(sd_dev, &sd, sizeof(SegDev), cudaMemcpyHostToDevice, st);
      LaunchTable lt;
      lt.n_segs = 1;
      lt.segs_host.push_back(sd);
      lt.segs_dev = sd_dev;
      tmp.eval[0] = lt;
      if (use_tc) rc = tc_build_tiles(&tmp.eval[0], mlp->in_dim);
      float* dummy = nullptr;   // no terms: nothing is written through `out`
      if (cudaMalloc(&dummy, sizeof(float) * (size_t)(tmp.P + 1)) != cudaSuccess) rc = fail(PINN_E_ALLOC, "scratch");
      if (rc == PINN_OK) rc = use_tc ? run_tc(&tmp, params_dev, dummy, st, false) : run_layered(&tmp, params_dev, dummy, st, false);
      cudaStreamSynchronize(st);
      if (dummy) cudaFree(dummy);
    }
    if (sd_dev) cudaFree(sd_dev);
    if (tmp.act) cudaFree(tmp.act);
    if (tmp.wt) cudaFree(tmp.wt);
    if (tmp.wimg) cudaFree(tmp.wimg);
    if (tmp.tc_rows) cudaFree(tmp.tc_rows);
    tmp.tc_rows = nullptr;
    tmp.act = tmp.wt = tmp.wimg = nullptr;
    if (tmp.eval[0].tiles_dev) cudaFree(tmp.eval[0].tiles_dev);
    tmp.eval[0].tiles_dev = nullptr;
    tmp.eval[0].segs_dev = nullptr;
    return rc;
  }
  int num_sms = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device));
  CUDA_TRY(cudaFuncSetAttribute((const void*)k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, k.smem_bytes));
  SegDev sd;
  memset(&sd, 0, sizeof(sd));
  sd.pts = points_dev;
  sd.y_out = y_dev;
  sd.n = n;
  sd.n_chunks = (int)((n + kChunk - 1) / kChunk);
  SegDev* sd_dev = nullptr;
  CUDA_TRY(cudaMallocAsync(&sd_dev, sizeof(SegDev), st));
  CUDA_TRY(cudaMemcpyAsync(sd_dev, &sd, sizeof(SegDev), cudaMemcpyHostToDevice, st));
  int grid = sd.n_chunks < num_sms ? sd.n_chunks : num_sms;
  const int aligned = ((uintptr_t)params_dev & 15u) == 0;
  k.fn<<<grid, k.nw * 32, k.smem_bytes, st>>>(params_dev, sd_dev, 1, sd.n_chunks, nullptr, 0, 0, aligned);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaFreeAsync(sd_dev, st));
  return PINN_OK;
}

// ------------------------------------------------------------------------------------------------
// NCCL through the library already loaded in the process (torch bundles libnccl.so.2)
// ------------------------------------------------------------------------------------------------
struct nccl_uid { char internal[128]; };
typedef int (*nccl_get_uid_fn)(nccl_uid*);
typedef int (*nccl_init_rank_fn)(void**, int, nccl_uid, int);
typedef int (*nccl_destroy_fn)(void*);
typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*nccl_errstr_fn)(int);

static struct {
  void* h = nullptr;
  nccl_get_uid_fn get_uid = nullptr;
  nccl_init_rank_fn init_rank = nullptr;
  nccl_destroy_fn destroy = nullptr;
  nccl_allreduce_fn allreduce = nullptr;
  nccl_errstr_fn errstr = nullptr;
} g_nccl;

static int nccl_load() {
  if (g_nccl.h) return PINN_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail(PINN_E_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
  g_nccl.get_uid = (nccl_get_uid_fn)dlsym(h, "ncclGetUniqueId");
  g_nccl.init_rank = (nccl_init_rank_fn)dlsym(h, "ncclCommInitRank");
  g_nccl.destroy = (nccl_destroy_fn)dlsym(h, "ncclCommDestroy");
  g_nccl.allreduce = (nccl_allreduce_fn)dlsym(h, "ncclAllReduce");
  g_nccl.errstr = (nccl_errstr_fn)dlsym(h, "ncclGetErrorString");
  if (!g_nccl.get_uid || !g_nccl.init_rank || !g_nccl.destroy || !g_nccl.allreduce)
    return fail(PINN_E_NCCL, "libnccl is missing a required symbol");
  g_nccl.h = h;
  return PINN_OK;
}

static int nccl_fail(const char* what, int rc) {
  return fail(PINN_E_NCCL, "%s: %s", what, g_nccl.errstr ? g_nccl.errstr(rc) : "nccl error");
}

extern "C" int pinn_nccl_unique_id(void* id128_out) {
  if (!id128_out) return fail(PINN_E_INVALID, "null argument");
  int rc = nccl_load();
  if (rc) return rc;
  nccl_uid id;
  int n = g_nccl.get_uid(&id);
  if (n) return nccl_fail("ncclGetUniqueId", n);
  memcpy(id128_out, &id, sizeof(id));
  return PINN_OK;
}

extern "C" int pinn_comm_create(const void* id128, int32_t world, int32_t rank, int32_t device, void** comm_out) {
  if (!id128 || !comm_out) return fail(PINN_E_INVALID, "null argument");
  int rc = nccl_load();
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(device));
  nccl_uid id;
  memcpy(&id, id128, sizeof(id));
  void* comm = nullptr;
  int n = g_nccl.init_rank(&comm, world, id, rank);
  if (n) return nccl_fail("ncclCommInitRank", n);
  *comm_out = comm;
  return PINN_OK;
}

extern "C" int pinn_comm_destroy(void* comm) {
  if (!comm) return PINN_OK;
  int rc = nccl_load();
  if (rc) return rc;
  int n = g_nccl.destroy(comm);
  return n ? nccl_fail("ncclCommDestroy", n) : PINN_OK;
}

extern "C" int pinn_allreduce_sum(void* comm, float* buf_dev, int64_t count, void* stream) {
  if (!comm || !buf_dev || count < 0) return fail(PINN_E_INVALID, "bad argument");
  int rc = nccl_load();
  if (rc) return rc;
  int n = g_nccl.allreduce(buf_dev, buf_dev, (size_t)count, /*ncclFloat32*/ 7, /*ncclSum*/ 0, comm, (cudaStream_t)stream);
  return n ? nccl_fail("ncclAllReduce", n) : PINN_OK;
}

// ------------------------------------------------------------------------------------------------
// Adam (Keras 2.7 update rule) on the flat vector
// ------------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ theta, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float step_size, float b1, float b2, float eps) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i];
  const float mi = fmaf(b1, m[i], (1.f - b1) * gi);
  const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
  m[i] = mi;
  v[i] = vi;
  theta[i] -= step_size * mi / (sqrtf(vi) + eps);
}

extern "C" int pinn_adam_step(float* params_dev, const float* grad_dev, float* m_dev, float* v_dev, int64_t count,
                              float lr, float beta1, float beta2, float eps, int64_t step, void* stream) {
  if (!params_dev || !grad_dev || !m_dev || !v_dev || count < 0 || step < 1) return fail(PINN_E_INVALID, "bad argument");
  if (count == 0) return PINN_OK;
  const double t = (double)step;
  const float step_size = (float)(lr * sqrt(1.0 - pow((double)beta2, t)) / (1.0 - pow((double)beta1, t)));
  adam_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(params_dev, grad_dev, m_dev, v_dev,
                                                                                  count, step_size, beta1, beta2, eps);
  CUDA_TRY(cudaGetLastError());
  return PINN_OK;
}

// graph-capturable variant: the step number lives on the device
__global__ void adam_dev_kernel(float* __restrict__ theta, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                                int64_t* __restrict__ step) {
  // *step: number of steps taken (low word; < 2^32).  The high word is this launch's arrival counter, zero between launches: the
  // last CTA to finish advances the step number -- no second launch for the increment.
  unsigned* words = reinterpret_cast<unsigned*>(step);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const double t = (double)words[0] + 1.0;
  if (i < n) {
    const float step_size = (float)((double)lr * sqrt(1.0 - pow((double)b2, t)) / (1.0 - pow((double)b1, t)));
    const float gi = g[i];
    const float mi = fmaf(b1, m[i], (1.f - b1) * gi);
    const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    theta[i] -= step_size * mi / (sqrtf(vi) + eps);
  }
  __syncthreads();                                        // every thread of this CTA has read the step number
  if (threadIdx.x == 0 && atomicAdd(words + 1, 1u) == gridDim.x - 1) {
    words[1] = 0u;
    words[0] += 1u;
  }
}
__global__ void bump_step_kernel(int64_t* step) { *step += 1; }

extern "C" int pinn_adam_step_dev(float* params_dev, const float* grad_dev, float* m_dev, float* v_dev, int64_t count,
                                  float lr, float beta1, float beta2, float eps, int64_t* step_dev, void* stream) {
  if (!params_dev || !grad_dev || !m_dev || !v_dev || !step_dev || count < 0) return fail(PINN_E_INVALID, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (count > 0)
    adam_dev_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(params_dev, grad_dev, m_dev, v_dev, count, lr, beta1, beta2,
                                                                     eps, step_dev);
  else
    bump_step_kernel<<<1, 1, 0, st>>>(step_dev);
  CUDA_TRY(cudaGetLastError());
  return PINN_OK;
}

// ------------------------------------------------------------------------------------------------
// One-shot all-reduce over NVLink peer memory (csrc/p2p.cuh): the ranks of one node map each other's receive blocks via CUDA IPC.
// ------------------------------------------------------------------------------------------------
struct pinn_p2p {
  int world = 0, rank = 0, device = 0;
  int64_t cap = 0;
  void* block = nullptr;                 // [2][world][cap] floats, then [2][kMaxWorld] flags
  void* opened[pinn::p2p::kMaxWorld] = {};
  pinn::p2p::Peers peers = {};
  uint32_t* epoch = nullptr;
  int* status = nullptr;
  bool connected = false;
};

static size_t p2p_data_bytes(int world, int64_t cap) { return ((size_t)2 * world * cap * sizeof(float) + 255) & ~(size_t)255; }

extern "C" int pinn_p2p_create(int32_t world, int32_t rank, int32_t device, int64_t max_count, void* handle64_out, void** ctx_out) {
  if (!handle64_out || !ctx_out || world < 2 || world > pinn::p2p::kMaxWorld || rank < 0 || rank >= world || max_count <= 0)
    return fail(PINN_E_INVALID, "pinn_p2p_create: bad argument (2 <= world <= %d)", pinn::p2p::kMaxWorld);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  DeviceGuard on_device(device);
  if (on_device.err != cudaSuccess) return fail(PINN_E_CUDA, "cannot switch to device %d", device);
  pinn_p2p* c = new (std::nothrow) pinn_p2p();
  if (!c) return fail(PINN_E_ALLOC, "host allocation");
  c->world = world; c->rank = rank; c->device = device;
  c->cap = (max_count + 63) & ~(int64_t)63;
  const size_t bytes = p2p_data_bytes(world, c->cap) + 2 * pinn::p2p::kMaxWorld * sizeof(uint32_t);
  cudaIpcMemHandle_t h;
  if (cudaMalloc(&c->block, bytes) != cudaSuccess || cudaMemset(c->block, 0, bytes) != cudaSuccess ||
      cudaMalloc(&c->epoch, sizeof(uint32_t)) != cudaSuccess || cudaMemset(c->epoch, 0, sizeof(uint32_t)) != cudaSuccess ||
      cudaMalloc(&c->status, sizeof(int)) != cudaSuccess || cudaMemset(c->status, 0, sizeof(int)) != cudaSuccess ||
      cudaIpcGetMemHandle(&h, c->block) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
    const char* e = cudaGetErrorString(cudaGetLastError());
    if (c->block) cudaFree(c->block);
    if (c->epoch) cudaFree(c->epoch);
    if (c->status) cudaFree(c->status);
    delete c;
    return fail(PINN_E_CUDA, "pinn_p2p_create: %s", e);
  }
  memcpy(handle64_out, &h, sizeof(h));
  *ctx_out = c;
  return PINN_OK;
}

extern "C" int pinn_p2p_connect(void* ctx, const void* handles) {
  pinn_p2p* c = (pinn_p2p*)ctx;
  if (!c || !handles) return fail(PINN_E_INVALID, "null argument");
  DeviceGuard on_device(c->device);
  const size_t data_bytes = p2p_data_bytes(c->world, c->cap);
  for (int r = 0; r < c->world; ++r) {
    void* base = c->block;
    if (r != c->rank) {
      cudaIpcMemHandle_t h;
      memcpy(&h, (const char*)handles + (size_t)r * sizeof(h), sizeof(h));
      cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(PINN_E_CUDA, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
      }
      c->opened[r] = base;
    }
    c->peers.data[r] = (float*)base;
    c->peers.flags[r] = (uint32_t*)((char*)base + data_bytes);
  }
  c->connected = true;
  return PINN_OK;
}

extern "C" int pinn_p2p_allreduce_sum(void* ctx, float* buf_dev, int64_t count, void* stream) {
  pinn_p2p* c = (pinn_p2p*)ctx;
  if (!c || !buf_dev || count < 0 || count > c->cap) return fail(PINN_E_INVALID, "pinn_p2p_allreduce_sum: bad argument");
  if (!c->connected) return fail(PINN_E_INVALID, "pinn_p2p_allreduce_sum: not connected");
  pinn::p2p::allreduce_oneshot_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(c->peers, c->world, c->rank, c->cap, buf_dev, (int)count, c->epoch,
                                                                         c->status);
  CUDA_TRY(cudaGetLastError());
  return PINN_OK;
}

/* number of waits that timed out so far (0 in a healthy run); synchronises the device */
extern "C" int pinn_p2p_status(void* ctx, int32_t* timeouts_out) {
  pinn_p2p* c = (pinn_p2p*)ctx;
  if (!c || !timeouts_out) return fail(PINN_E_INVALID, "null argument");
  DeviceGuard on_device(c->device);
  int v = 0;
  CUDA_TRY(cudaMemcpy(&v, c->status, sizeof(int), cudaMemcpyDeviceToHost));
  *timeouts_out = v;
  return PINN_OK;
}

extern "C" int pinn_p2p_destroy(void* ctx) {
  pinn_p2p* c = (pinn_p2p*)ctx;
  if (!c) return PINN_OK;
  DeviceGuard on_device(c->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < c->world; ++r)
    if (c->opened[r]) cudaIpcCloseMemHandle(c->opened[r]);
  if (c->block) cudaFree(c->block);
  if (c->epoch) cudaFree(c->epoch);
  if (c->status) cudaFree(c->status);
  delete c;
  return PINN_OK;
}

// ------------------------------------------------------------------------------------------------
// BFGS round: quasi-Newton algebra on the device (csrc/bfgs.cuh).  All pointers are device pointers of the caller; everything
// is enqueued on the caller's stream; `scal` is a device array of PINN_BFGS_SCALARS doubles the caller reads back.
// ------------------------------------------------------------------------------------------------
extern "C" int pinn_bfgs_identity(double* H, int64_t n, void* stream) {
  if (!H || n <= 0) return fail(PINN_E_INVALID, "bad argument");
  const int64_t total = n * n;
  pinn::bfgs::bfgs_identity_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(H, n);
  CUDA_TRY(cudaGetLastError());
  return PINN_OK;
}

extern "C" int pinn_bfgs_trial(const double* x, const double* p, double alpha, double* xt, float* theta, int64_t n, void* stream) {
  if (!x || !p || !xt || !theta || n <= 0) return fail(PINN_E_INVALID, "bad argument");
  pinn::bfgs::bfgs_trial_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, p, alpha, xt, theta, n);
  CUDA_TRY(cudaGetLastError());
  return PINN_OK;
}

extern "C" int pinn_bfgs_trial_dev(const double* x, const double* p, const double* alpha_dev, double* xt, float* theta, int64_t n,
                                   void* stream) {
  if (!x || !p || !alpha_dev || !xt || !theta || n <= 0) return fail(PINN_E_INVALID, "bad argument");
  pinn::bfgs::bfgs_trial_dev_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, p, alpha_dev, xt, theta, n);
  CUDA_TRY(cudaGetLastError());
  return PINN_OK;
}

extern "C" int pinn_bfgs_eval(const float* out, const double* coef, const int32_t* kind, int32_t n_terms, const double* p, double* gt,
                              double* scal, int64_t n, void* stream) {
  if (!out || !p || !gt || !scal || n <= 0 || n_terms < 0 || (n_terms > 0 && (!coef || !kind))) return fail(PINN_E_INVALID, "bad argument");
  pinn::bfgs::bfgs_eval_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(out, coef, kind, n_terms, p, gt, scal, n);
  CUDA_TRY(cudaGetLastError());
  return PINN_OK;
}

extern "C" int pinn_bfgs_direction(const double* H, const double* g, double* p, double* scal, int64_t n, void* stream) {
  if (!H || !g || !p || !scal || n <= 0) return fail(PINN_E_INVALID, "bad argument");
  pinn::bfgs::bfgs_matvec_kernel<<<(unsigned)((n + 7) / 8), 256, 0, (cudaStream_t)stream>>>(H, g, p, -1.0, n);
  pinn::bfgs::bfgs_slope_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(g, p, scal, n);
  CUDA_TRY(cudaGetLastError());
  return PINN_OK;
}

extern "C" int pinn_bfgs_accept_update(double* H, double* x, double* g, const double* xt, const double* gt, double* s, double* y, double* u,
                                       double* p, double* scal, int32_t update_h, int64_t n, void* stream) {
  if (!H || !x || !g || !xt || !gt || !s || !y || !u || !p || !scal || n <= 0) return fail(PINN_E_INVALID, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  pinn::bfgs::bfgs_accept_kernel<<<1, 1024, 0, st>>>(x, g, xt, gt, s, y, scal, n);
  if (update_h) {
    pinn::bfgs::bfgs_matvec_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(H, y, u, 1.0, n);
    pinn::bfgs::bfgs_update_direction_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(H, s, y, u, g, p, scal, n);
    pinn::bfgs::bfgs_slope_kernel<<<1, 1024, 0, st>>>(g, p, scal, n);
  }
  CUDA_TRY(cudaGetLastError());
  return PINN_OK;
}

#ifdef PINN_TC_PROFILE
// development build only (tools/tc_phase_profile.py): read and clear the per-phase cycle counters of fused_tc_kernel
extern "C" int pinn_tc_profile_read(unsigned long long* host_out) {
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpyFromSymbol(host_out, pinn::ftc::g_tc_prof, sizeof(pinn::ftc::g_tc_prof)));
  static unsigned long long zero[160 * 20 * 16];
  CUDA_TRY(cudaMemcpyToSymbol(pinn::ftc::g_tc_prof, zero, sizeof(zero)));
  return PINN_OK;
}
extern "C" int pinn_tc_stage_read(long long* host_out) {
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpyFromSymbol(host_out, pinn::ftc::g_tc_stage, sizeof(pinn::ftc::g_tc_stage)));
  return PINN_OK;
}
// raw clock64 stamps of CTA 0, tiles 2..9 of its sequence (tools/tc_trace.py)
extern "C" int pinn_tc_trace_read(long long* host_out) {
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpyFromSymbol(host_out, pinn::ftc::g_tc_trace, sizeof(pinn::ftc::g_tc_trace)));
  return PINN_OK;
}
#endif
