/*
 * pinnstep.h -- C ABI of libpinnstep.so, the B200 (sm_100a) PINN loss-step library.
 *
 * The reference (giuliamesc/PINNs_Fluid_Dynamics) has no FFI: its training step sits behind
 * nisaba's Python API and is handed over as Python closures over TensorFlow ops.  This header is
 * the boundary a maintainer would bind instead; every entry point names the reference code it
 * replaces (paths relative to the reference repository root).
 *
 *   model(x) + every inner gradient(tape, ., x)     Examples/Cavity_Steady/cavity_steady.py:159-188,
 *                                                   Examples/Cavity_Unsteady/cavity_unsteady.py:169-199
 *   PDE_MASS / PDE_MOM / dir_loss / neu_loss / PDE  cavity_steady.py:159-200, poiseuille_flow.py:173-209,
 *                                                   colliding_flow.py:160-184, poisson_misto.py:62-80
 *   ns.LossMeanSquares (mean of squared roots)      cavity_steady.py:204,212-225
 *   ns.OptimizationProblem (sum_t w_t L_t, gradient
 *   w.r.t. model.variables)                         cavity_steady.py:242
 *
 * Conventions
 *   - plain C types only; all pointers named *_dev are DEVICE pointers owned by the caller and
 *     must stay valid for the lifetime of the plan (point coordinates and rhs arrays) or of the
 *     call (params / out).
 *   - every function returns 0 (PINN_OK) or a negative PINN_E_* code; the message of the last
 *     failure on the calling thread is available from pinn_last_error().  Nothing throws or
 *     aborts across this boundary.
 *   - work is enqueued asynchronously on the caller's stream (a cudaStream_t passed as void*);
 *     pinn_loss_and_grad / pinn_loss do not allocate and do not synchronise, so they can be
 *     captured in a CUDA graph.
 *   - a plan is bound to one device and is not thread-safe; one plan per rank.  The entry points that take a plan switch
 *     to its device for the call and restore the caller's current device; several plans on several devices may live in
 *     one process (per-device kernel attributes are set on first use on each device).
 *
 * Residual model.  Let J[o][c] be the Taylor jet of network output o at a point:
 *   c = 0               value
 *   c = 1 + i           d/dx_i              (i = 0..d-1, input column i)
 *   c = 1 + d           d2/dx_sx^2          (sx = d-2: first spatial column)
 *   c = 2 + d           d2/dx_sy^2          (sy = d-1: second spatial column)
 * A loss term evaluates, on every point n of its point set,
 *   r_n = sum_{o,c} coef[o][c] * J[o][c]
 *       + conv * ( J[0][0] * J[conv_k][1+sx] + J[1][0] * J[conv_k][1+sy] )
 *       - rhs_scale * rhs[n]
 * which covers every closure of the five in-scope scripts (DESIGN.md section 3 lists the
 * coefficient sets).  The term's loss value is  sum_n r_n^2 / (normalization * n_global)  and the
 * total loss is  sum_t weight_t * value_t  (cavity_steady.py:212-231, nisaba semantics).
 *
 * A term of kind PINN_TERM_ABS_MEAN restates  ns.Loss('PRESS_0', lambda: tf.abs(tf.reduce_mean(model(x)[:,2])))
 * (Examples/Colliding_Flow/colliding_flow_pressmean.py:176-179,196): its value is  |sum_n r_n| / (normalization *
 * n_global),  its output slot carries the raw  sum_n r_n,  and its gradient is  sign(sum_n r_n) * weight /
 * (normalization * n_global) * sum_n d r_n / d theta.  The sign is taken from a forward pre-pass over the term's point
 * set inside pinn_loss_and_grad, so such a set must live entirely on ONE rank (the host side keeps it on rank 0).
 */
#ifndef PINNSTEP_H
#define PINNSTEP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PINN_VERSION 102 /* 0.1.2: pinn_bfgs_* (quasi-Newton algebra of the BFGS round); 0.1.1: pinn_term_desc.kind, pinn_adam_step_dev */

#define PINN_MAX_OUT 4   /* network outputs (u, v, p) padded to 4 */
#define PINN_MAX_CH 6    /* value, d/dt, d/dx, d/dy, d2/dx2, d2/dy2 */
#define PINN_MAX_DIM 3
#define PINN_MAX_TERMS_PER_SET 8

enum pinn_term_kind {
  PINN_TERM_MEAN_SQUARES = 0, /* ns.LossMeanSquares */
  PINN_TERM_ABS_MEAN = 1      /* ns.Loss over |mean(residual)| (colliding_flow_pressmean.py:196) */
};

enum pinn_status {
  PINN_OK = 0,
  PINN_E_INVALID = -1,     /* bad argument / unsupported shape */
  PINN_E_ALLOC = -2,       /* device or host allocation failed */
  PINN_E_CUDA = -3,        /* a CUDA runtime call failed */
  PINN_E_ARCH = -4,        /* device is not sm_100 */
  PINN_E_NCCL = -5,        /* NCCL missing or a collective failed */
  PINN_E_STATE = -6        /* call sequence error */
};

/* tanh MLP of the reference: Sequential([Dense(H,tanh)] * L + [Dense(O)]), cavity_steady.py:205-210.
 * Parameters are one flat fp32 vector in Keras variable order [K1,b1,...,K_{L+1},b_{L+1}], kernels
 * row-major [in,out] (y = x @ K + b). */
typedef struct pinn_mlp_desc {
  int32_t in_dim;      /* d: 2 (x,y) or 3 (t,x,y) */
  int32_t width;       /* H */
  int32_t n_hidden;    /* L */
  int32_t out_dim;     /* O: 1 (Poisson) or 3 (u,v,p) */
} pinn_mlp_desc;

/* One loss term = one ns.LossMeanSquares entry of the script's loss table. */
typedef struct pinn_term_desc {
  float coef[PINN_MAX_OUT][PINN_MAX_CH]; /* linear part, see header comment */
  float conv;                            /* coefficient of the convective product */
  int32_t conv_k;                        /* which velocity component is convected */
  float rhs_scale;                       /* multiplies rhs[n]; ignored when rhs_dev == NULL */
  const float* rhs_dev;                  /* [n_local] fp32 or NULL */
  double weight;                         /* ns.LossMeanSquares(weight=...) */
  double normalization;                  /* ns.LossMeanSquares(normalization=...) */
  int64_t n_global;                      /* number of roots of this term over ALL ranks */
  int32_t train;                         /* 1: contributes to the total loss and gradient;
                                            0: test loss (forward only, pinn_loss) */
  int32_t kind;                          /* PINN_TERM_MEAN_SQUARES (0) | PINN_TERM_ABS_MEAN (1) */
} pinn_term_desc;

/* One point set = one category of the scripts (PDE / one boundary edge / IC / Vel / Pres / Test). */
typedef struct pinn_pointset_desc {
  const float* points_dev;   /* [n_local, d] fp32 row-major */
  int64_t n_local;           /* rows on this rank (may be 0) */
  int32_t n_terms;           /* terms evaluated on this set, fused in one pass */
  int32_t deriv_order;       /* 0: value only; 1: + first derivatives; 2: + d2/dx2, d2/dy2 */
  pinn_term_desc terms[PINN_MAX_TERMS_PER_SET];
} pinn_pointset_desc;

typedef struct pinn_plan pinn_plan;

/* Library / device ------------------------------------------------------------------------- */
int pinn_version(void);
const char* pinn_last_error(void);

/* Plan ------------------------------------------------------------------------------------- */
/* Copies all descriptors; allocates the per-CTA partial workspace on `device`.  Term order in the
 * outputs is set order then term order within the set. */
int pinn_plan_create(const pinn_mlp_desc* mlp, const pinn_pointset_desc* sets, int32_t n_sets,
                     int32_t device, pinn_plan** out);
int pinn_plan_destroy(pinn_plan* plan);
int64_t pinn_plan_param_count(const pinn_plan* plan);
int32_t pinn_plan_term_count(const pinn_plan* plan);
size_t pinn_plan_workspace_bytes(const pinn_plan* plan);
/* Name of the engine chosen for the plan's networks: "fused_fp32" | "fused_tf32x3" | "layered_fp32" | "layered_tf32x3". */
const char* pinn_plan_engine(const pinn_plan* plan);
/* Kernel launches enqueued by the last pinn_loss_and_grad / pinn_loss call. */
int32_t pinn_plan_last_launch_count(const pinn_plan* plan);
/* Per-kernel device timing for benchmarks: when enabled, each fused kernel launch of the plan is
 * bracketed by CUDA events on the caller's stream (this makes the call non graph-capturable).
 * pinn_plan_kernel_time_ms returns the duration of the last launch of the kernel that handled the
 * point sets of derivative order `deriv_order` (2 = collocation kernel); it synchronises on that event. */
int pinn_plan_enable_timing(pinn_plan* plan, int32_t on);
int pinn_plan_kernel_time_ms(pinn_plan* plan, int32_t deriv_order, float* ms_out);
/* Replace the rhs array of one term (e.g. re-drawn noise) without rebuilding the plan. */
int pinn_plan_set_rhs(pinn_plan* plan, int32_t set_index, int32_t term_index, const float* rhs_dev);

/* Hot path --------------------------------------------------------------------------------- */
/* out_dev: [P + T] fp32.  out[0..P) = d/dtheta of  sum_{t in train} weight_t/(normalization_t*n_global_t) * sum_n r_n^2
 * over the LOCAL points; out[P + t] = local sum_n r_n^2 of term t (train and test terms alike; test
 * terms are only filled by pinn_loss).  With every rank's out summed (pinn_allreduce_sum or any SUM
 * all-reduce) the result is the exact global gradient and the global sums of squares.
 * Replaces: nisaba's per-epoch  loss = sum_t w_t*LMS_t();  grads = tape.gradient(loss, variables). */
int pinn_loss_and_grad(pinn_plan* plan, const float* params_dev, float* out_dev, void* stream);
/* Forward only (no gradient): fills out[P + t] for ALL terms, leaves out[0..P) untouched.
 * Replaces: the evaluation of `losses_test` at every log point (cavity_steady.py:233-235). */
int pinn_loss(pinn_plan* plan, const float* params_dev, float* out_dev, void* stream);
/* Forward-only network evaluation, values only: y[n, o] = model(x)[n, o].
 * Replaces: model(grid) in the scripts' post-processing (cavity_steady.py:254-268).
 * Fused engines (H <= 32): asynchronous on `stream` (stream-ordered scratch), graph-capturable.  Layered engines (H = 64,
 * 128): the EXCEPTION to the conventions above -- the call allocates a temporary activation workspace sized for n points,
 * and returns after the work has completed and the workspace is freed (post-processing is not a hot path; the training
 * entry points never do this). */
int pinn_forward(const pinn_mlp_desc* mlp, const float* params_dev, const float* points_dev,
                 int64_t n, float* y_dev, int32_t device, void* stream);

/* Multi-GPU -------------------------------------------------------------------------------- */
/* The only cross-GPU dependency of a step is the SUM of the [P+T] vector.  These wrap the NCCL
 * that is already loaded in the process (dlopen "libnccl.so.2"); a torch.distributed all_reduce on
 * the same buffer is equivalent. */
int pinn_nccl_unique_id(void* id128_out);                          /* 128-byte ncclUniqueId */
int pinn_comm_create(const void* id128, int32_t world, int32_t rank, int32_t device, void** comm_out);
int pinn_comm_destroy(void* comm);
int pinn_allreduce_sum(void* comm, float* buf_dev, int64_t count, void* stream);
/* The same SUM as ONE kernel over NVLink peer memory (ranks of one node, 2 <= world <= 8; csrc/p2p.cuh): every rank stores its
 * vector into a receive block of every peer (mapped through CUDA IPC), publishes a step flag, waits for its peers' flags and sums
 * the `world` vectors in rank order -- bit-identical on all ranks, graph-capturable, ~3x lower latency than ncclAllReduce at 9 KB.
 * pinn_p2p_create returns this rank's 64-byte IPC handle; the caller gathers the handles of all ranks (rank order, e.g. with a
 * torch.distributed all_gather) and passes the world x 64 bytes to pinn_p2p_connect.  Every rank must call
 * pinn_p2p_allreduce_sum the same number of times.  pinn_p2p_status reports waits that timed out (0 in a healthy run). */
int pinn_p2p_create(int32_t world, int32_t rank, int32_t device, int64_t max_count, void* handle64_out, void** ctx_out);
int pinn_p2p_connect(void* ctx, const void* handles_world_x_64);
int pinn_p2p_allreduce_sum(void* ctx, float* buf_dev, int64_t count, void* stream);
int pinn_p2p_status(void* ctx, int32_t* timeouts_out);
int pinn_p2p_destroy(void* ctx);

/* Optimiser helpers on the flat vectors (rows (f) rank 1 of SURVEY.md section 8) ------------ */
/* Keras Adam: m,v update with bias correction folded in the step size, epsilon outside the sqrt
 * (tf.keras.optimizers.Adam(learning_rate=1e-2), cavity_steady.py:246). step is 1-based. */
int pinn_adam_step(float* params_dev, const float* grad_dev, float* m_dev, float* v_dev, int64_t count,
                   float lr, float beta1, float beta2, float eps, int64_t step, void* stream);
/* Same update with the 1-based step number kept in device memory (*step_dev holds the number of steps already
 * taken and is incremented by the call): no host-side value changes between steps, so a whole training step
 * (pinn_loss_and_grad + this) can be captured once in a CUDA graph and replayed. */
int pinn_adam_step_dev(float* params_dev, const float* grad_dev, float* m_dev, float* v_dev, int64_t count,
                       float lr, float beta1, float beta2, float eps, int64_t* step_dev, void* stream);

/* BFGS round (ns.minimize(pb, 'scipy', 'BFGS', epochs), cavity_steady.py:247; nisaba hands it to scipy.optimize.minimize):
 * the quasi-Newton algebra on the device.  State, all float64 device arrays of the caller: the iterate x [n], its gradient
 * g [n], the direction p [n], trial point / gradient xt, gt [n], s, y, u [n], the dense inverse Hessian H [n x n] row-major,
 * and scal [PINN_BFGS_SCALARS]: [0] phi(alpha), [1] phi'(alpha) = gt . p, [2] |gt|_inf, [3] y . s, [4] g . p of the new
 * direction, [5] |p|_2.  Everything is enqueued on `stream`; the host reads back scal only. */
#define PINN_BFGS_SCALARS 8
int pinn_bfgs_identity(double* H_dev, int64_t n, void* stream);                       /* H = I (SciPy's start) */
/* xt = x + alpha p, theta = float(xt): the parameters of the next loss step */
int pinn_bfgs_trial(const double* x_dev, const double* p_dev, double alpha, double* xt_dev, float* theta_dev, int64_t n, void* stream);
/* the same with alpha read from device memory at execution time: lets the caller capture trial point + loss step + evaluation in
 * ONE CUDA graph and replay it for every step length of the line search */
int pinn_bfgs_trial_dev(const double* x_dev, const double* p_dev, const double* alpha_dev, double* xt_dev, float* theta_dev, int64_t n,
                        void* stream);
/* after pinn_loss_and_grad (+ the SUM over ranks) into out [n + n_terms]: gt = double(gradient), scal[0] = sum_t coef[t] *
 * (kind[t] ? |out[n+t]| : out[n+t]) (the total loss: coef = weight / (normalization N_global), kind 1 = |mean| term),
 * scal[1] = gt . p, scal[2] = |gt|_inf */
int pinn_bfgs_eval(const float* out_dev, const double* coef_dev, const int32_t* kind_dev, int32_t n_terms, const double* p_dev,
                   double* gt_dev, double* scal_dev, int64_t n, void* stream);
/* p = -H g, scal[4] = g . p, scal[5] = |p|_2 */
int pinn_bfgs_direction(const double* H_dev, const double* g_dev, double* p_dev, double* scal_dev, int64_t n, void* stream);
/* accept the trial point: s = xt - x, y = gt - g, x = xt, g = gt, scal[3] = y . s; with update_h != 0 also
 * H <- H - rho (s u^T + u s^T) + (rho^2 y.u + rho) s s^T, u = H y, rho = 1 / (y . s) (1000 if y . s == 0, like SciPy), and the next
 * direction p = -H g with scal[4], scal[5] -- one pass over H for the update and the direction together */
int pinn_bfgs_accept_update(double* H_dev, double* x_dev, double* g_dev, const double* xt_dev, const double* gt_dev, double* s_dev,
                            double* y_dev, double* u_dev, double* p_dev, double* scal_dev, int32_t update_h, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PINNSTEP_H */
