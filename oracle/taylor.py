"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never import this from the product package.

Float64 numpy restatement of the algorithm the CUDA kernels implement: Taylor-mode jets
(value, d/dx_i, d2/dx2, d2/dy2) pushed through the tanh MLP layer by layer, the
residual forms of include/pinnstep.h, and the hand-derived reverse sweep through the jets
(SURVEY.md D.2).  It evaluates a ``CompiledProblem`` (the product's host-side compilation of a loss
table) and returns the same [P + T] vector ``pinn_loss_and_grad`` produces.

Its role is to tie the two halves of the parity argument together:
  nested reverse-mode (oracle/reference_step.py, the reference's formulation)
      == Taylor-mode + hand reverse (this file)            -- tests/test_oracle_taylor.py, 1e-10
      == CUDA FP32 kernels                                  -- tests/test_parity_gpu.py, 1e-5 / 1e-4
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def unpack(theta: np.ndarray, d: int, H: int, L: int, O: int):
    """Keras order [K1, b1, ..., K_{L+1}, b_{L+1}], kernels [in, out]."""
    Ks, bs, off = [], [], 0
    sizes = [d] + [H] * L + [O]
    for i in range(L + 1):
        n = sizes[i] * sizes[i + 1]
        Ks.append(theta[off:off + n].reshape(sizes[i], sizes[i + 1])); off += n
        bs.append(theta[off:off + sizes[i + 1]]); off += sizes[i + 1]
    assert off == theta.size
    return Ks, bs


def n_channels(d: int, order: int) -> int:
    return 1 if order == 0 else (1 + d if order == 1 else 3 + d)


def forward_jets(Ks, bs, x: np.ndarray, order: int):
    """Returns (J [C, N, O], stash) with stash[l] = (a, zj) per hidden layer: a [N,H] the tanh value,
    zj [C, N, H] the pre-activation jet."""
    N, d = x.shape
    C = n_channels(d, order)
    sx, sy = d - 2, d - 1
    L = len(Ks) - 1
    H = Ks[0].shape[1]
    zj = np.zeros((C, N, H))
    zj[0] = x @ Ks[0] + bs[0]
    if order >= 1:
        for i in range(d):
            zj[1 + i] = Ks[0][i][None, :]
    stash = []
    aj = None
    for l in range(L):
        if l > 0:
            zj = (aj.reshape(C * N, H) @ Ks[l]).reshape(C, N, -1)
            zj[0] += bs[l]
        a = np.tanh(zj[0])
        s = 1.0 - a * a
        q = -2.0 * a * s
        aj = np.zeros_like(zj)
        aj[0] = a
        if order >= 1:
            for i in range(d):
                aj[1 + i] = s * zj[1 + i]
        if order >= 2:
            aj[1 + d] = q * zj[1 + sx] ** 2 + s * zj[1 + d]
            aj[2 + d] = q * zj[1 + sy] ** 2 + s * zj[2 + d]
        stash.append((a, zj, aj))
    J = (aj.reshape(C * N, H) @ Ks[L]).reshape(C, N, -1)
    J[0] += bs[L]
    return J, stash


def _coefs(term):
    """float64 coefficients straight from the form (the device receives them rounded to fp32; that
    6e-8 relative difference is inside the stated parity tolerance)."""
    m = np.zeros((4, 6))
    for (o, c), v in term.form.coef.items():
        m[o, c] = v
    return m, float(term.form.conv), float(term.form.rhs_scale)


def residual(term, J: np.ndarray, d: int, rhs) -> np.ndarray:
    """term: CompiledTerm.  J [C, N, O]."""
    C = J.shape[0]
    sx, sy = d - 2, d - 1
    m, conv, rhs_scale = _coefs(term)
    r = np.zeros(J.shape[1])
    for o in range(J.shape[2]):
        for c in range(C):
            if m[o, c] != 0.0:
                r += m[o, c] * J[c, :, o]
    if conv != 0.0:
        k = term.form.conv_k
        r += conv * (J[0, :, 0] * J[1 + sx, :, k] + J[0, :, 1] * J[1 + sy, :, k])
    if rhs is not None:
        r -= rhs_scale * rhs
    return r


def residual_adjoint(term, J, rbar, d: int) -> np.ndarray:
    sx, sy = d - 2, d - 1
    m, cv, _ = _coefs(term)
    Jb = np.zeros_like(J)
    for o in range(J.shape[2]):
        for c in range(J.shape[0]):
            if m[o, c] != 0.0:
                Jb[c, :, o] += m[o, c] * rbar
    if cv != 0.0:
        k = term.form.conv_k
        Jb[0, :, 0] += cv * rbar * J[1 + sx, :, k]
        Jb[1 + sx, :, k] += cv * rbar * J[0, :, 0]
        Jb[0, :, 1] += cv * rbar * J[1 + sy, :, k]
        Jb[1 + sy, :, k] += cv * rbar * J[0, :, 1]
    return Jb


def backward_jets(Ks, x, order, stash, Jbar):
    """Reverse sweep: returns flat gradient in Keras order."""
    N, d = x.shape
    sx, sy = d - 2, d - 1
    L = len(Ks) - 1
    gK = [np.zeros_like(K) for K in Ks]
    gb = [np.zeros(K.shape[1]) for K in Ks]
    # output layer
    aj = stash[L - 1][2]
    C = aj.shape[0]
    H = aj.shape[2]
    gK[L] = aj.reshape(C * N, H).T @ Jbar.reshape(C * N, -1)
    gb[L] = Jbar[0].sum(axis=0)
    ab = (Jbar.reshape(C * N, -1) @ Ks[L].T).reshape(C, N, H)
    for l in range(L - 1, -1, -1):
        a, zj, _ = stash[l]
        s = 1.0 - a * a
        q = -2.0 * a * s
        qp = -2.0 * s * (1.0 - 3.0 * a * a)
        zb = np.zeros_like(ab)
        zb[0] = s * ab[0]
        if order >= 1:
            acc = np.zeros_like(a)
            for i in range(d):
                zb[1 + i] = s * ab[1 + i]
                acc += zj[1 + i] * ab[1 + i]
            zb[0] += q * acc
        if order >= 2:
            zb[1 + d] = s * ab[1 + d]
            zb[2 + d] = s * ab[2 + d]
            zb[1 + sx] += 2.0 * q * zj[1 + sx] * ab[1 + d]
            zb[1 + sy] += 2.0 * q * zj[1 + sy] * ab[2 + d]
            zb[0] += q * (zj[1 + d] * ab[1 + d] + zj[2 + d] * ab[2 + d])
            zb[0] += qp * (zj[1 + sx] ** 2 * ab[1 + d] + zj[1 + sy] ** 2 * ab[2 + d])
        if l > 0:
            aprev = stash[l - 1][2]
            gK[l] = aprev.reshape(C * N, H).T @ zb.reshape(C * N, H)
            gb[l] = zb[0].sum(axis=0)
            ab = (zb.reshape(C * N, H) @ Ks[l].T).reshape(C, N, H)
        else:
            gK[0] = x.T @ zb[0]
            if order >= 1:
                for i in range(d):
                    gK[0][i] += zb[1 + i].sum(axis=0)
            gb[0] = zb[0].sum(axis=0)
    return np.concatenate([np.concatenate([k.reshape(-1), b]) for k, b in zip(gK, gb)])


def loss_and_grad(cp, theta: np.ndarray, with_grad: bool = True, include_test: bool = False, chunk: int = 32768) -> np.ndarray:
    """[P + T] float64: local gradient of sum_train w/(nu N_global) sum r^2, then local sum r^2 per
    term in TABLE order (CompiledTerm.out_index).  Point sets are walked in chunks of ``chunk`` points (every term
    is a sum over points), so the BASELINE sizes (1 M / 4 M collocation points) fit host memory."""
    d, H, L, O = cp.mlp
    theta = np.asarray(theta, dtype=np.float64)
    Ks, bs = unpack(theta, d, H, L, O)
    out = np.zeros(cp.n_params + max(1, cp.n_out_terms))
    for cs in cp.sets:
        if cs.n_local == 0:
            continue
        # |mean| terms (ns.Loss over |mean(roots)|): the adjoint needs the sign of the sum over the WHOLE set first
        signs = {}
        if with_grad and any(t.abs_mean and t.train for t in cs.terms):
            sums = {id(t): 0.0 for t in cs.terms if t.abs_mean}
            for a in range(cs.start, cs.stop, chunk):
                b = min(cs.stop, a + chunk)
                J, _ = forward_jets(Ks, bs, cs.pointset.points[a:b].astype(np.float64), cs.deriv_order)
                for t in cs.terms:
                    if t.abs_mean:
                        rhs = t.form.rhs_array()
                        sums[id(t)] += float(np.sum(residual(t, J, d, None if rhs is None else rhs[a:b].astype(np.float64))))
            signs = {k: (-1.0 if v < 0 else 1.0) for k, v in sums.items()}
        for a in range(cs.start, cs.stop, chunk):
            b = min(cs.stop, a + chunk)
            x = cs.pointset.points[a:b].astype(np.float64)
            J, stash = forward_jets(Ks, bs, x, cs.deriv_order)
            Jbar = np.zeros_like(J)
            for t in cs.terms:
                if not t.train and not include_test:
                    continue
                rhs = t.form.rhs_array()
                rhs = None if rhs is None else rhs[a:b].astype(np.float64)
                r = residual(t, J, d, rhs)
                if t.abs_mean:     # slot = sum r, adjoint = sign(sum r) w / (nu N)
                    out[cp.n_params + t.out_index] += float(np.sum(r))
                    if t.train and with_grad:
                        Jbar += residual_adjoint(t, J, np.full_like(r, signs[id(t)] * t.weight / (t.normalization * t.n_global)), d)
                    continue
                out[cp.n_params + t.out_index] += float(np.sum(r * r))
                if t.train and with_grad:
                    scale = 2.0 * t.weight / (t.normalization * t.n_global)
                    Jbar += residual_adjoint(t, J, scale * r, d)
            if with_grad:
                out[:cp.n_params] += backward_jets(Ks, x, cs.deriv_order, stash, Jbar)
    return out


class TaylorEngine:
    """Drop-in for ``CudaPlan`` in CPU-only tests of the host logic (sharding, all-reduce, history):
    injected through ``OptimizationProblem(engine_factory=...)``.  Test use only."""

    def __init__(self, cp):
        import torch
        self.cp, self.P = cp, cp.n_params
        self.torch = torch
        self.engine = "oracle_taylor_fp64"

    def loss_and_grad(self, params_flat):
        o = loss_and_grad(self.cp, params_flat.detach().cpu().double().numpy())
        return self.torch.as_tensor(o, dtype=self.torch.float64)

    def loss_only(self, params_flat):
        o = loss_and_grad(self.cp, params_flat.detach().cpu().double().numpy(), with_grad=False, include_test=True)
        return self.torch.as_tensor(o, dtype=self.torch.float64)

    def to_table_order(self, out):
        return out[self.P:self.P + max(1, self.cp.n_out_terms)]

    def last_launch_count(self) -> int:
        return 0
