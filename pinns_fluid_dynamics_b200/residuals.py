"""Declarative residual specifications -- what replaces the reference's Python closures.

The reference hands its loss terms to nisaba as opaque closures over TensorFlow ops
(``LMS('PDE_MOMU', lambda: PDE_MOM(0), weight=1e0)``, cavity_steady.py:212-225).  A fused CUDA
kernel cannot run a closure, so each closure of the five in-scope scripts has a builder here that
returns a ``ResidualForm``: the coefficients of

    r_n = sum_{o,c} coef[o][c] * J[o][c](x_n)
        + conv * ( J[0][val] * J[conv_k][dx] + J[1][val] * J[conv_k][dy] )
        - rhs_scale * rhs[n]

over the output jets J (include/pinnstep.h).  The builders keep the reference's argument meaning
(component indices, edge targets, norm_vel / norm_pre scaling) and its quirks as
explicit, named defaults.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np

MAX_OUT, MAX_CH = 4, 6


class PointSet:
    """A fixed set of points ``[n, d]`` (one category: PDE, one boundary edge, IC, Vel, Pres, Test).

    Terms built on the SAME PointSet object are evaluated in one fused pass over the set -- the
    reference re-runs ``model(x)`` for every term (3 forwards on the PDE set per step,
    cavity_steady.py:212-214)."""

    _next_id = 0

    def __init__(self, points, name: str = ""):
        pts = np.ascontiguousarray(np.asarray(points, dtype=np.float32))
        if pts.ndim != 2:
            raise ValueError("points must be [n, d]")
        self.points = pts
        self.name = name or f"set{PointSet._next_id}"
        self.uid = PointSet._next_id
        PointSet._next_id += 1

    @property
    def n(self) -> int:
        return self.points.shape[0]

    @property
    def dim(self) -> int:
        return self.points.shape[1]


def ch_val() -> int:
    return 0


def ch_d(i: int) -> int:
    """first derivative w.r.t. input column i"""
    return 1 + i


def ch_dd(dim: int, which: int) -> int:
    """second derivative w.r.t. spatial column ``which`` (0 -> x, 1 -> y)"""
    return 1 + dim + which


def spatial_cols(dim: int) -> Tuple[int, int]:
    """(x, y) input columns: (0, 1) steady; (1, 2) when t is column 0 (cavity_unsteady.py:95)."""
    return dim - 2, dim - 1


@dataclass
class ResidualForm:
    pointset: PointSet
    coef: Dict[Tuple[int, int], float] = field(default_factory=dict)   # (output, channel) -> coefficient
    conv: float = 0.0
    conv_k: int = 0
    rhs: Optional[np.ndarray] = None
    rhs_scale: float = 1.0
    identically_zero: bool = False      # quirk Q1: logs 0.0, contributes no gradient
    reduction: str = "mean_squares"     # "mean_squares": mean(r^2) (ns.LossMeanSquares); "abs_mean": |mean(r)| (ns.Loss)
    source: str = ""                    # reference closure this form restates (file:line)

    def deriv_order(self) -> int:
        dim = self.pointset.dim
        order = 0
        chans = [c for (_, c), v in self.coef.items() if v != 0.0]
        if self.conv != 0.0:
            chans += [ch_d(spatial_cols(dim)[0])]
        for c in chans:
            if c >= 1 + dim:
                order = max(order, 2)
            elif c >= 1:
                order = max(order, 1)
        return order

    def coef_matrix(self) -> np.ndarray:
        m = np.zeros((MAX_OUT, MAX_CH), dtype=np.float32)
        for (o, c), v in self.coef.items():
            m[o, c] = v
        return m

    def rhs_array(self) -> Optional[np.ndarray]:
        if self.rhs is None:
            return None
        r = np.ascontiguousarray(np.asarray(self.rhs, dtype=np.float32).reshape(-1))
        if r.shape[0] != self.pointset.n:
            raise ValueError(f"rhs has {r.shape[0]} entries for a point set of {self.pointset.n}")
        return r


# --------------------------------------------------------------------------------------------
# builders, one per reference closure
# --------------------------------------------------------------------------------------------

def dirichlet(pointset: PointSet, component: int, rhs=None) -> ResidualForm:
    """``dir_loss(points, component, rhs)``: model(points)[:, component] - rhs
    (cavity_steady.py:192-194; its lambdas BC_D, IN_C, fit_velocity, fit_pressure, exact_value
    :196-200).  Also Poisson's ``BC_D`` / ``lambda: model(x_BC)`` (poisson_misto.py:69-73,
    poisson.py:67) with rhs=None and ``lambda: model(x_test) - u_test`` (poisson.py:69)."""
    return ResidualForm(pointset, {(component, ch_val()): 1.0}, rhs=rhs, rhs_scale=1.0,
                        source="cavity_steady.py:192-194")


def mass(pointset: PointSet, in_tape: bool = True, scale: float = 1.0) -> ResidualForm:
    """``PDE_MASS``: d_x N_0 + d_y N_1 on the normalised outputs (cavity_steady.py:159-166,
    cavity_unsteady.py:169-176).  ``in_tape=False`` restates Colliding_Flow / Poiseuille_Flow, where
    ``divergence(tape, u_vect, x, dim)`` runs after the tape closed (colliding_flow.py:160-165,
    poiseuille_flow.py:173-178) and the term is identically zero (quirk Q1).  ``scale`` = vel_max restates the pressmean
    variant's in-tape divergence of ``model(x)[:, 0:2] * vel_max`` (colliding_flow_pressmean.py:140-145)."""
    sx, sy = spatial_cols(pointset.dim)
    return ResidualForm(pointset, {(0, ch_d(sx)): float(scale), (1, ch_d(sy)): float(scale)},
                        identically_zero=not in_tape, source="cavity_steady.py:159-166")


def momentum(pointset: PointSet, k: int, norm_vel: float, norm_pre: float, *,
             conv_scale: float, visc_xx: float, visc_yy: float, time_derivative: bool = False) -> ResidualForm:
    """``PDE_MOM(k)`` with u_eq = norm_vel*N_k, p = norm_pre*N_2, kappa = 1/max(norm_pre, norm_vel):

        kappa * [ d_t u_eq (if time_derivative) + visc_xx * d_xx u_eq + visc_yy * d_yy u_eq + d_k p
                  + conv_scale * (N_0 * d_x u_eq + N_1 * d_y u_eq) ]

    Cavity_Steady   (cavity_steady.py:168-188):   conv_scale=norm_vel, visc_xx=+1, visc_yy=-1 (quirk Q2)
    Cavity_Unsteady (cavity_unsteady.py:178-199): conv_scale=norm_vel, visc_xx=visc_yy=-1, time_derivative
    Colliding_Flow  (colliding_flow.py:167-184):  conv_scale=1 (quirk Q4), visc_xx=visc_yy=-1
    Poiseuille_Flow (poiseuille_flow.py:180-197): conv_scale=rho (Q4), visc_xx=visc_yy=-mu
    """
    dim = pointset.dim
    sx, sy = spatial_cols(dim)
    kappa = 1.0 / max(norm_pre, norm_vel)
    coef = {
        (k, ch_dd(dim, 0)): kappa * visc_xx * norm_vel,
        (k, ch_dd(dim, 1)): kappa * visc_yy * norm_vel,
        (2, ch_d((sx, sy)[k])): kappa * norm_pre,
    }
    if time_derivative:
        coef[(k, ch_d(0))] = kappa * norm_vel
    return ResidualForm(pointset, coef, conv=kappa * conv_scale * norm_vel, conv_k=k,
                        source="cavity_steady.py:168-188")


def neumann(pointset: PointSet, k: int, j: int, rhs, norm_vel: float, norm_pre: float, mu: float) -> ResidualForm:
    """``neu_loss(x, k, j, rhs)``: kappa * (mu * d_j (norm_vel*N_k) - norm_pre*N_2*[j==k] - rhs)
    (poiseuille_flow.py:199-209)."""
    kappa = 1.0 / max(norm_pre, norm_vel)
    coef = {(k, ch_d(j)): kappa * mu * norm_vel}
    if j == k:
        coef[(2, ch_val())] = -kappa * norm_pre
    return ResidualForm(pointset, coef, rhs=rhs, rhs_scale=kappa, source="poiseuille_flow.py:199-209")


def outflow_stress(pointset: PointSet, k: int, normal, rhs, norm_vel: float, norm_pre: float, ni: float,
                   in_tape: bool = False) -> ResidualForm:
    """Coronary ``neu_loss(edge, k, rhs)``: ni * grad(norm_vel*N_k) . n - norm_pre*N_2 * n[k] - rhs with the script's
    un-normalised normals n = (2, 1) on OUT1 and (1, 0) on OUT2 (coronary_flow_steady.py:197-211).  There the model
    is called AFTER the tape closed (:205-209), so ``gradient`` returns zeros and only the pressure part survives --
    ``in_tape=False`` restates that (for n[k] = 0 the residual is the constant -rhs: the flat BCN_v_OUT2 log of
    Test_Case_#123); ``in_tape=True`` is the intended traction condition."""
    coef = {(2, ch_val()): -norm_pre * float(normal[k])}
    if in_tape:
        for j in (0, 1):
            if normal[j] != 0:
                coef[(k, ch_d(j))] = ni * norm_vel * float(normal[j])
    return ResidualForm(pointset, coef, rhs=rhs, rhs_scale=1.0, source="coronary_flow_steady.py:197-211")


def stokes_momentum(pointset: PointSet, k: int, vel_max: float, p_max: float, forcing=None) -> ResidualForm:
    """Pressmean variant ``PDE_MOM(x, k, force)``: -laplacian(vel_max*N_k) + d_k (p_max*N_2) - force(x)
    (colliding_flow_pressmean.py:147-159; no convective term, no normalisation constant)."""
    dim = pointset.dim
    return ResidualForm(pointset, {(k, ch_dd(dim, 0)): -float(vel_max), (k, ch_dd(dim, 1)): -float(vel_max),
                                   (2, ch_d(k)): float(p_max)}, rhs=forcing, rhs_scale=1.0,
                        source="colliding_flow_pressmean.py:147-159")


def mean_value(pointset: PointSet, component: int) -> ResidualForm:
    """``PRESS_0(x)``: |mean(model(x)[:, component])| -- the scalar handed to ``ns.Loss``
    (colliding_flow_pressmean.py:176-179,196).  The roots are N_component(x_n); the reduction is |mean|."""
    return ResidualForm(pointset, {(component, ch_val()): 1.0}, reduction="abs_mean",
                        source="colliding_flow_pressmean.py:176-179")


def poisson_pde(pointset: PointSet, forcing) -> ResidualForm:
    """``PDE``: -laplacian(u) - f (poisson.py:58-63, poisson_misto.py:62-67)."""
    dim = pointset.dim
    return ResidualForm(pointset, {(0, ch_dd(dim, 0)): -1.0, (0, ch_dd(dim, 1)): -1.0},
                        rhs=forcing, rhs_scale=1.0, source="poisson.py:58-63")


def normal_derivative(pointset: PointSet, component: int, direction: int, rhs) -> ResidualForm:
    """Poisson ``BC_N``: d_x u - g (poisson_misto.py:75-80)."""
    return ResidualForm(pointset, {(component, ch_d(direction)): 1.0}, rhs=rhs, rhs_scale=1.0,
                        source="poisson_misto.py:75-80")
