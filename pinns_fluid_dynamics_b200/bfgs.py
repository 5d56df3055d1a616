"""BFGS round of the scripts (``ns.minimize(pb, 'scipy', 'BFGS', num_epochs=epochs)``, cavity_steady.py:247) with the
quasi-Newton algebra on the device.

nisaba hands the problem to ``scipy.optimize.minimize(method='BFGS')``.  SciPy 1.x updates the dense inverse Hessian
as ``Hk = A1 @ (Hk @ A2) + rho s s^T`` -- two P x P x P products, 24.5 GFLOP per iteration for the 2307-parameter
network: measured 133 ms per iteration on the GPU box's host next to a 1.9 ms loss step (tools/optimizer_rounds.py).
This module keeps SciPy's algorithm -- same start (H0 = I, first step ~ 1), the same strong-Wolfe line search (SciPy's
own ``_line_search_wolfe12``: MINPACK dcsrch, then the Nocedal-Wright zoom), the same stopping rules and the same
curvature safeguard -- and replaces only the update by its algebraically identical rank-2 form

    u = H y,   H <- H - rho (s u^T + u s^T) + (rho^2 y.u + rho) s s^T,        p = -H g,

three passes over a float64 P x P tensor that lives on the GPU (PyTorch as the tensor shell).
``tests/test_bfgs.py`` checks the iterates against ``scipy.optimize.minimize`` itself.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np
import torch

try:  # SciPy's own line search keeps the step decisions identical to the reference's driver
    from scipy.optimize._optimize import _line_search_wolfe12, _LineSearchError
except Exception:  # pragma: no cover - very old / very new SciPy
    _line_search_wolfe12 = None

    class _LineSearchError(RuntimeError):
        pass


class _Memo:
    """f and g of the same point are one device step (SciPy's ScalarFunction does the same caching)."""

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


class BfgsResult(dict):
    __getattr__ = dict.get


def minimize_bfgs(fun: Callable[[np.ndarray], Tuple[float, np.ndarray]], x0: np.ndarray, maxiter: Optional[int] = None,
                  callback: Optional[Callable[[np.ndarray], None]] = None, gtol: float = 1e-5, c1: float = 1e-4,
                  c2: float = 0.9, xrtol: float = 0.0, device="cuda") -> BfgsResult:
    """``scipy.optimize._optimize._minimize_bfgs`` (norm = inf, jac = True) with the inverse Hessian on ``device``.
    ``fun(x) -> (f, grad)`` takes and returns host float64 arrays (the PINN step computes in FP32 behind it)."""
    if _line_search_wolfe12 is None:
        raise RuntimeError("scipy.optimize._optimize._line_search_wolfe12 is not importable in this SciPy")
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
        try:
            alpha_k, _, _, old_fval, old_old_fval, gfkp1 = _line_search_wolfe12(
                memo.f_of, memo.g_of, xk, pk, gfk, old_fval, old_old_fval, amin=1e-100, amax=1e100, c1=c1, c2=c2)
        except _LineSearchError:
            warnflag = 2
            break
        sk = alpha_k * pk
        xk = xk + sk
        if gfkp1 is None:
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
