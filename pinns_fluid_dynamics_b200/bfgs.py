"""BFGS round of the scripts (``ns.minimize(pb, 'scipy', 'BFGS', num_epochs=epochs)``, cavity_steady.py:247).

nisaba hands the problem to ``scipy.optimize.minimize(method='BFGS')``.  SciPy 1.x updates the dense inverse Hessian
as ``Hk = A1 @ (Hk @ A2) + rho s s^T`` -- two P x P x P products, 24.5 GFLOP per iteration for the 2307-parameter
network: measured 133 ms per iteration on the GPU box's host next to a 0.8 ms loss step.  This module keeps SciPy's
algorithm -- same start (H0 = I, first step ~ 1), the same strong-Wolfe line search (More-Thuente, then the
Nocedal-Wright zoom: linesearch.py), the same stopping rules and the same curvature safeguard -- with the algebraically
identical rank-2 form of the update

    u = H y,   H <- H - rho (s u^T + u s^T) + (rho^2 y.u + rho) s s^T,        p = -H g.

Two drivers:
  * ``minimize_bfgs_device(pb, ...)`` -- the product path.  Iterate, gradient, direction and the float64 inverse Hessian
    live in device memory; hand-written kernels (csrc/bfgs.cuh through the C ABI: ``pinn_bfgs_*``) form the trial point,
    turn the loss step's output into (phi, phi', |g|_inf), and apply update + next direction in ONE pass over H.  The
    host runs only the scalar logic of the line search: it reads back three doubles per evaluation.
  * ``minimize_bfgs(fun, x0, ...)`` -- the same algorithm on host arrays for any ``fun(x) -> (f, grad)``;
    tests/test_bfgs.py checks its iterates against ``scipy.optimize.minimize`` itself.
When the dense inverse Hessian does not fit the device (P = 116 483 of the 8x128 network: 108 GB) the device driver
switches to the limited-memory two-loop recursion (m = 20 pairs) on the same line search.
"""
from __future__ import annotations

import ctypes as C
import os
import time
from typing import Callable, Optional, Tuple

import numpy as np
import torch

from .linesearch import first_step, more_thuente


class LineSearchError(RuntimeError):
    pass


class BfgsResult(dict):
    __getattr__ = dict.get


def _wolfe2_fallback(f_of, g_of, xk, pk, gfk, old_fval, old_old_fval, c1, c2):
    """Nocedal-Wright bracketing / zoom search: SciPy's public ``line_search`` (what its BFGS falls back to)."""
    import warnings
    import scipy.optimize
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        alpha, _, _, fval, old, gval = scipy.optimize.line_search(f_of, g_of, xk, pk, gfk, old_fval, old_old_fval, c1=c1, c2=c2,
                                                                 amax=1e100)
    if alpha is None:
        raise LineSearchError("line search did not converge")
    return alpha, fval, old, gval


class _Memo:
    """f and g of the same point are one evaluation (SciPy's ScalarFunction does the same caching)."""

    def __init__(self, fun: Callable[[np.ndarray], Tuple[float, np.ndarray]]):
        self.fun, self.x, self.f, self.g, self.nfev = fun, None, None, None, 0

    def _eval(self, x):
        if self.x is None or not np.array_equal(x, self.x):
            self.x = np.array(x, dtype=np.float64, copy=True)
            f, g = self.fun(self.x)
            self.f, self.g = float(f), np.asarray(g, dtype=np.float64)
            self.nfev += 1

    def f_of(self, x):
        self._eval(x)
        return self.f

    def g_of(self, x):
        self._eval(x)
        return self.g


def minimize_bfgs(fun: Callable[[np.ndarray], Tuple[float, np.ndarray]], x0: np.ndarray, maxiter: Optional[int] = None,
                  callback: Optional[Callable[[np.ndarray], None]] = None, gtol: float = 1e-5, c1: float = 1e-4,
                  c2: float = 0.9, xrtol: float = 0.0, device="cpu") -> BfgsResult:
    """``scipy.optimize.minimize(method='BFGS', jac=True)`` (norm = inf) with the O(P^2) form of the update.
    ``fun(x) -> (f, grad)`` takes and returns host float64 arrays; ``device`` places the inverse Hessian."""
    dev = torch.device(device)
    x0 = np.asarray(x0, dtype=np.float64).reshape(-1)
    N = x0.shape[0]
    maxiter = N * 200 if maxiter is None else int(maxiter)
    memo = _Memo(fun)
    old_fval = memo.f_of(x0)
    gfk = memo.g_of(x0)
    H = torch.eye(N, dtype=torch.float64, device=dev)
    old_old_fval = old_fval + np.linalg.norm(gfk) / 2      # first step ~ 1 (SciPy)
    xk, k, warnflag = x0, 0, 0
    gnorm = np.abs(gfk).max() if N else 0.0
    while gnorm > gtol and k < maxiter:
        g_dev = torch.as_tensor(gfk, dtype=torch.float64, device=dev)
        pk = (-torch.mv(H, g_dev)).cpu().numpy()
        derphi0 = float(np.dot(gfk, pk))
        trial = {}

        def phi(a):
            xa = xk + a * pk
            fa, ga = memo.f_of(xa), memo.g_of(xa)
            trial["g"] = ga
            return fa, float(np.dot(ga, pk))

        alpha_k, fval, _ = more_thuente(phi, old_fval, derphi0, first_step(old_fval, old_old_fval, derphi0), ftol=c1, gtol=c2)
        if alpha_k is not None:
            old_old_fval, old_fval, gfkp1 = old_fval, fval, trial["g"]
        else:
            try:
                alpha_k, fval, old_old_fval, gfkp1 = _wolfe2_fallback(memo.f_of, memo.g_of, xk, pk, gfk, old_fval, old_old_fval, c1, c2)
                old_fval = fval
            except LineSearchError:
                warnflag = 2
                break
        sk = alpha_k * pk
        xk = xk + sk
        if gfkp1 is None or np.ndim(gfkp1) == 0:
            gfkp1 = memo.g_of(xk)
        yk = gfkp1 - gfk
        gfk = gfkp1
        k += 1
        if callback is not None:
            callback(xk)
        gnorm = np.abs(gfk).max()
        if gnorm <= gtol:
            break
        if alpha_k * np.linalg.norm(pk) <= xrtol * (xrtol + np.linalg.norm(xk)):
            break
        if not np.isfinite(old_fval):
            warnflag = 2
            break
        rhok_inv = float(np.dot(yk, sk))
        rhok = 1000.0 if rhok_inv == 0.0 else 1.0 / rhok_inv
        s = torch.as_tensor(sk, dtype=torch.float64, device=dev)
        y = torch.as_tensor(yk, dtype=torch.float64, device=dev)
        u = torch.mv(H, y)
        a = float(torch.dot(y, u))
        H.addr_(s, u, alpha=-rhok)
        H.addr_(u, s, alpha=-rhok)
        H.addr_(s, s, alpha=rhok * rhok * a + rhok)
    if warnflag == 0 and k >= maxiter:
        warnflag = 1
    return BfgsResult(x=xk, fun=old_fval, jac=gfk, nit=k, nfev=memo.nfev, status=warnflag, success=(warnflag == 0),
                      hess_inv=H)


# ------------------------------------------------------------------------------------------------
# device driver
# ------------------------------------------------------------------------------------------------

def dense_hessian_fits(n: int, device) -> bool:
    """The float64 P x P inverse Hessian (plus slack) against the free memory of ``device``."""
    free, _ = torch.cuda.mem_get_info(device)
    return n * n * 8 + (256 << 20) < 0.8 * free


def minimize_bfgs_device(pb, maxiter: int, callback: Optional[Callable[[], None]] = None, gtol: float = 1e-5, c1: float = 1e-4,
                         c2: float = 0.9, history_m: int = 20) -> BfgsResult:
    """The BFGS round of ``pb`` (an ``OptimizationProblem`` on a CUDA plan) with every vector and the inverse Hessian on the
    device.  ``pb.flat`` holds the accepted iterate (FP32) whenever ``callback()`` runs and on return."""
    from . import _capi
    plan, lib = pb.plan, pb.plan.lib
    dev = pb.flat.device
    n = int(pb.compiled.n_params)
    f64 = dict(dtype=torch.float64, device=dev)
    x = pb.flat.detach().double().clone()
    g, p, xt, gt, s, y, u = (torch.zeros(n, **f64) for _ in range(7))
    scal = torch.zeros(8, **f64)
    scal_host = torch.zeros(8, dtype=torch.float64, pin_memory=True)
    # PINN_BFGS_DENSE=0 forces the limited-memory recursion (tests; the 8x128 network takes it by itself)
    dense = dense_hessian_fits(n, dev) and os.environ.get("PINN_BFGS_DENSE", "1") != "0"
    H = torch.empty((n, n), **f64) if dense else None
    S, Y = [], []                                  # limited-memory pairs (only when H does not fit)
    # total loss from the kernel-order term sums: phi = sum_t coef_t * (|v_t| for |mean| terms, else v_t)
    terms = [t for cs in pb.compiled.sets for t in cs.terms]
    coef = torch.tensor([(t.weight / (t.normalization * t.n_global) if (t.train and t.n_global) else (float("nan") if t.train else 0.0))
                         for t in terms], **f64)
    kind = torch.tensor([1 if t.abs_mean else 0 for t in terms], dtype=torch.int32, device=dev)
    vp = C.c_void_p

    def stream():
        return vp(torch.cuda.current_stream(dev).cuda_stream)

    def ptr(t):
        return vp(t.data_ptr())

    def read_scalars():
        scal_host.copy_(scal, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return scal_host

    nfev = [0]
    # One evaluation = trial point + loss step (+ SUM over ranks) + evaluation kernel + read-back of 8 doubles.  The same
    # launch sequence every time (the step length travels through a pinned host scalar), so it is captured once in a CUDA
    # graph and replayed: an evaluation costs one graph launch and one stream synchronisation instead of ~8 launches
    # (at 10 k collocation points the launches, not the kernels, set the pace).  PINN_BFGS_GRAPH=0 keeps it eager.
    use_graph = os.environ.get("PINN_BFGS_GRAPH", "1") != "0" and not getattr(plan, "timing_enabled", False)
    alpha_host = torch.zeros(1, dtype=torch.float64, pin_memory=True)
    alpha_np, scal_np = alpha_host.numpy(), scal_host.numpy()      # views of the pinned buffers: scalar access without tensor overhead
    alpha_dev = torch.zeros(1, **f64)
    graphs = {"eval": None, "accept": None}

    def enqueue_eval():
        alpha_dev.copy_(alpha_host, non_blocking=True)
        _capi.check(lib.pinn_bfgs_trial_dev(ptr(x), ptr(p), ptr(alpha_dev), ptr(xt), ptr(pb.flat), n, stream()), "pinn_bfgs_trial_dev")
        out = pb._reduce(plan.loss_and_grad(pb.flat))
        _capi.check(lib.pinn_bfgs_eval(ptr(out), ptr(coef), ptr(kind), len(terms), ptr(p), ptr(gt), ptr(scal), n, stream()),
                    "pinn_bfgs_eval")
        scal_host.copy_(scal, non_blocking=True)

    def enqueue_accept():
        _capi.check(lib.pinn_bfgs_accept_update(ptr(H), ptr(x), ptr(g), ptr(xt), ptr(gt), ptr(s), ptr(y), ptr(u), ptr(p), ptr(scal),
                                                1, n, stream()), "pinn_bfgs_accept_update")
        pb.flat.copy_(x)                               # FP32 copy of the accepted iterate
        scal_host.copy_(scal, non_blocking=True)

    spent = {"eval": 0.0, "accept": 0.0, "n_eval": 0, "n_accept": 0}

    def run(name, enqueue):
        t_begin = time.perf_counter()
        try:
            return _run(name, enqueue)
        finally:
            spent[name] += time.perf_counter() - t_begin
            spent["n_" + name] += 1

    def _run(name, enqueue):
        """replay the captured sequence (captured on its second use: the first one runs eagerly and creates whatever the
        step creates lazily -- NCCL communicator, kernel attributes), then wait for the scalars"""
        if use_graph and graphs[name] is None:
            graphs[name] = False                       # next call captures
            enqueue()
        elif use_graph and graphs[name] is False:
            torch.cuda.synchronize(dev)
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, capture_error_mode="thread_local" if pb.world > 1 else "global"):
                enqueue()
            graphs[name] = gr
            gr.replay()
        elif use_graph:
            graphs[name].replay()
        else:
            enqueue()
        torch.cuda.current_stream(dev).synchronize()
        return scal_np

    def evaluate(alpha: float):
        """phi(alpha), phi'(alpha), |g|_inf at x + alpha p"""
        alpha_np[0] = alpha
        h = run("eval", enqueue_eval)
        nfev[0] += 1
        return float(h[0]), float(h[1]), float(h[2])

    def direction_lbfgs():
        """two-loop recursion (Nocedal & Wright Alg. 7.4) on device vectors: p = -H_k g with H_0 = gamma I"""
        q = g.clone()
        al = []
        for sv, yv in zip(reversed(S), reversed(Y)):
            rho = 1.0 / torch.dot(yv, sv)
            a = rho * torch.dot(sv, q)
            q.add_(yv, alpha=-float(a))
            al.append((a, rho))
        if S:
            q.mul_(float(torch.dot(S[-1], Y[-1]) / torch.dot(Y[-1], Y[-1])))
        for (a, rho), sv, yv in zip(reversed(al), S, Y):
            b = rho * torch.dot(yv, q)
            q.add_(sv, alpha=float(a - b))
        p.copy_(q).neg_()
        scal[4] = torch.dot(g, p)

    # start: f(x0), g(x0) through the same kernels with p = 0
    phi0, _, gnorm = evaluate(0.0)
    g.copy_(gt)
    if dense:
        _capi.check(lib.pinn_bfgs_identity(ptr(H), n, stream()), "pinn_bfgs_identity")
        _capi.check(lib.pinn_bfgs_direction(ptr(H), ptr(g), ptr(p), ptr(scal), n, stream()), "pinn_bfgs_direction")
    else:
        direction_lbfgs()
    old_fval = phi0
    old_old_fval = old_fval + float(torch.linalg.vector_norm(g)) / 2      # first step ~ 1 (SciPy)
    k, warnflag = 0, 0
    derphi0 = float(read_scalars()[4])
    while gnorm > gtol and k < maxiter:
        last = {}

        def phi(a):
            fa, da, gn = evaluate(a)
            last["gn"] = gn
            return fa, da

        alpha_k, fval, _ = more_thuente(phi, old_fval, derphi0, first_step(old_fval, old_old_fval, derphi0), ftol=c1, gtol=c2)
        if alpha_k is None:
            # rare: More-Thuente stopped on a warning -> SciPy's public zoom search on host copies of the vectors
            xk, pk, gk = x.cpu().numpy(), p.cpu().numpy(), g.cpu().numpy()
            try:
                alpha_k, fval, old_old_fval2, _ = _wolfe2_fallback(lambda z: pb.evaluate_host(z)[0], lambda z: pb.evaluate_host(z)[1],
                                                                   xk, pk, gk, old_fval, old_old_fval, c1, c2)
            except LineSearchError:
                warnflag = 2
                break
            fval2, _, gn = evaluate(alpha_k)          # leaves x_t, g_t of the accepted step on the device
            last["gn"] = gn
            fval = fval2
        old_old_fval, old_fval = old_fval, fval
        gnorm = last["gn"]
        k += 1
        # the accepted point is x_t; pb.flat already holds float(x_t) unless the last trial was another alpha
        update = gnorm > gtol and np.isfinite(old_fval)
        if dense and update:
            derphi0 = float(run("accept", enqueue_accept)[4])
        elif dense:
            _capi.check(lib.pinn_bfgs_accept_update(ptr(H), ptr(x), ptr(g), ptr(xt), ptr(gt), ptr(s), ptr(y), ptr(u), ptr(p), ptr(scal),
                                                    0, n, stream()), "pinn_bfgs_accept_update")
            pb.flat.copy_(x)
        else:
            s.copy_(xt).sub_(x)
            y.copy_(gt).sub_(g)
            x.copy_(xt)
            g.copy_(gt)
            if update:
                if float(torch.dot(y, s)) > 0.0:
                    S.append(s.clone()); Y.append(y.clone())
                    if len(S) > history_m:
                        S.pop(0); Y.pop(0)
                direction_lbfgs()
                derphi0 = float(read_scalars()[4])
            pb.flat.copy_(x)                           # FP32 copy of the accepted iterate
        if callback is not None:
            callback()
        if not update:
            if not np.isfinite(old_fval):
                warnflag = 2
            break
    if warnflag == 0 and k >= maxiter:
        warnflag = 1
    pb.flat.copy_(x)
    return BfgsResult(x=x, fun=old_fval, nit=k, nfev=nfev[0], status=warnflag, success=(warnflag == 0), hess_inv=H, seconds=spent,
                      algebra="dense float64 inverse Hessian on the device" if dense else f"L-BFGS two-loop, m = {history_m}")
