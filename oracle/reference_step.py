"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never import this from the product package.

Reference-faithful CPU restatement (torch float64, nested reverse-mode autodiff) of the loss tables
of the in-scope example scripts (five BASELINE configs + Coronary_Flow).  One ``model(x)`` forward PER LOSS TERM and one
``gradient(tape, ., x)`` sweep per call in the script -- including the duplicated sweeps -- so that
this is also a fair stand-in for timing the reference's step on the CPU (bench.py cpu_baseline,
kind "port").  PARITY UNPINNED at the bit level (TensorFlow / nisaba cannot run here), see oracle/nisaba_like.py;
what the reference's saved artefacts pin numerically (trained-weight replays, sol_pinn.h5) is in tests/test_oracle_pins.py
and tests/test_coronary.py.

Each builder takes the ``ProblemData`` arrays (plain numpy) and the Keras-ordered weights and
returns ``OptimizationProblem(variables, losses, losses_test)`` exactly as the script's last lines do.
Quirks Q1-Q4 are reproduced by default: they are the reference's behaviour.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch

from .nisaba_like import (DTYPE, GradientTape, KerasMLP, Loss, LossMeanSquares, OptimizationProblem,
                          divergence_vector, gradient_scalar, laplacian_scalar)


def _t(a) -> torch.Tensor:
    return torch.as_tensor(np.asarray(a, dtype=np.float64), dtype=DTYPE)


def _watch(a: torch.Tensor) -> torch.Tensor:
    # tf.gather(dom_grid, idx) materialises a fresh tensor every call (cavity_steady.py:160)
    return a.detach().clone()


# --------------------------------------------------------------------------------------------
# Cavity_Steady  (Examples/Cavity_Steady/cavity_steady.py)
# --------------------------------------------------------------------------------------------

def cavity_steady(data, variables: Sequence[torch.Tensor]) -> OptimizationProblem:
    model = KerasMLP(variables)
    norm_vel, norm_pre = data.norm_vel, data.norm_pre
    x_pde = _t(data.x_pde)
    bnd_pts = {k: _t(v) for k, v in data.bnd_pts.items()}
    bnd_val = [{k: _t(v) for k, v in d.items()} for d in data.bnd_val]
    sol_noise = [_t(v) for v in data.sol_noise]
    sol_test = [_t(v) for v in data.sol_test]
    x_vel, x_pres, x_test = _t(data.x_vel), _t(data.x_pres), _t(data.x_test)
    gradient = gradient_scalar

    def PDE_MASS():  # cavity_steady.py:159-166
        x = _watch(x_pde)
        with GradientTape(persistent=True) as tape:
            tape.watch(x)
            u_vect = model(x)[:, 0:2]
            du_x = gradient(tape, u_vect[:, 0], x)[:, 0]
            dv_y = gradient(tape, u_vect[:, 1], x)[:, 1]
        return du_x + dv_y

    def PDE_MOM(k):  # cavity_steady.py:168-188
        x = _watch(x_pde)
        with GradientTape(persistent=True) as tape:
            tape.watch(x)
            u_vect = model(x)
            p = u_vect[:, 2] * norm_pre
            u_eq = u_vect[:, k] * norm_vel
            dp = gradient(tape, p, x)[:, k]
            du_x = gradient(tape, u_eq, x)[:, 0]
            du_y = gradient(tape, u_eq, x)[:, 1]     # same sweep recomputed (Q6)
            du_xx = gradient(tape, du_x, x)[:, 0]
            du_yy = gradient(tape, du_y, x)[:, 1]
            conv1 = torch.mul(norm_vel * u_vect[:, 0], du_x)
            conv2 = torch.mul(norm_vel * u_vect[:, 1], du_y)
            unnormed_lhs = du_xx - du_yy + dp + conv1 + conv2   # sign quirk Q2 (:185)
            norm_const = 1 / max(norm_pre, norm_vel)
        return unnormed_lhs * norm_const

    def dir_loss(points, component, rhs):  # :192-194
        uk = model(points)[:, component]
        return uk - rhs

    BC_D = lambda edge, component: dir_loss(bnd_pts[edge], component, bnd_val[component][edge])
    fit_velocity = lambda component: dir_loss(x_vel, component, sol_noise[component])
    fit_pressure = lambda: dir_loss(x_pres, 2, sol_noise[2])
    exact_value = lambda component: dir_loss(x_test, component, sol_test[component])

    LMS = LossMeanSquares
    o = data.options
    losses: List[LossMeanSquares] = []
    if o.use_collloss:
        losses += [LMS('PDE_MASS', lambda: PDE_MASS(), weight=1e1),
                   LMS('PDE_MOMU', lambda: PDE_MOM(0), weight=1e0),
                   LMS('PDE_MOMV', lambda: PDE_MOM(1), weight=1e0)]
    if o.use_boundary:
        losses += [LMS('BCD_u_x0', lambda: BC_D("SX", 0)), LMS('BCD_v_x0', lambda: BC_D("SX", 1)),
                   LMS('BCD_u_x1', lambda: BC_D("DX", 0)), LMS('BCD_v_x1', lambda: BC_D("DX", 1)),
                   LMS('BCD_u_y0', lambda: BC_D("BOT", 0)), LMS('BCD_v_y0', lambda: BC_D("BOT", 1)),
                   LMS('BCD_u_y1', lambda: BC_D("TOP", 0)), LMS('BCD_v_y1', lambda: BC_D("TOP", 1))]
    # Q3: the booleans fit_velocity / fit_pressure are shadowed by the lambdas above (:198-199), so
    # `if fit_velocity:` (:230-231) is always true and the fit terms are always present.
    losses += [LMS('Fit_u', lambda: fit_velocity(0)), LMS('Fit_v', lambda: fit_velocity(1))]
    losses += [LMS('Fit_p', lambda: fit_pressure())]
    loss_test = [LMS('u_test', lambda: exact_value(0)), LMS('v_test', lambda: exact_value(1)),
                 LMS('p_test', lambda: exact_value(2))]
    return OptimizationProblem(model.variables, losses, loss_test)


# --------------------------------------------------------------------------------------------
# Cavity_Unsteady  (Examples/Cavity_Unsteady/cavity_unsteady.py)
# --------------------------------------------------------------------------------------------

def cavity_unsteady(data, variables: Sequence[torch.Tensor]) -> OptimizationProblem:
    model = KerasMLP(variables)
    norm_vel, norm_pre = data.norm_vel, data.norm_pre
    x_pde = _t(data.x_pde)
    bnd_pts = {k: _t(v) for k, v in data.bnd_pts.items()}
    bnd_val = [{k: _t(v) for k, v in d.items()} for d in data.bnd_val]
    sol_noise = [_t(v) for v in data.sol_noise]
    sol_test = [_t(v) for v in data.sol_test]
    x_vel, x_pres, x_test = _t(data.x_vel), _t(data.x_pres), _t(data.x_test)
    n_ic = data.bnd_pts["IC"].shape[0] if "IC" in data.bnd_pts else 0
    gradient = gradient_scalar

    def PDE_MASS():  # cavity_unsteady.py:169-176  (columns 1, 2 are x, y)
        x = _watch(x_pde)
        with GradientTape(persistent=True) as tape:
            tape.watch(x)
            u_vect = model(x)[:, 0:2]
            du_x = gradient(tape, u_vect[:, 0], x)[:, 1]
            dv_y = gradient(tape, u_vect[:, 1], x)[:, 2]
        return du_x + dv_y

    def PDE_MOM(k):  # cavity_unsteady.py:178-199
        x = _watch(x_pde)
        with GradientTape(persistent=True) as tape:
            tape.watch(x)
            u_vect = model(x)
            p = u_vect[:, 2] * norm_pre
            u_eq = u_vect[:, k] * norm_vel
            dp = gradient(tape, p, x)[:, k + 1]
            du_t = gradient(tape, u_eq, x)[:, 0]
            du_x = gradient(tape, u_eq, x)[:, 1]
            du_y = gradient(tape, u_eq, x)[:, 2]
            du_xx = gradient(tape, du_x, x)[:, 1]
            du_yy = gradient(tape, du_y, x)[:, 2]
            conv1 = torch.mul(norm_vel * u_vect[:, 0], du_x)
            conv2 = torch.mul(norm_vel * u_vect[:, 1], du_y)
            unnormed_lhs = du_t - du_xx - du_yy + dp + conv1 + conv2
            norm_const = 1 / max(norm_pre, norm_vel)
        return unnormed_lhs * norm_const

    def dir_loss(points, component, rhs):
        uk = model(points)[:, component]
        return uk - rhs

    BC_D = lambda edge, component: dir_loss(bnd_pts[edge], component, bnd_val[component][edge])
    IN_C = lambda component: dir_loss(bnd_pts["IC"], component, torch.zeros(n_ic, dtype=DTYPE))
    fit_velocity = lambda component: dir_loss(x_vel, component, sol_noise[component])
    fit_pressure = lambda: dir_loss(x_pres, 2, sol_noise[2])
    exact_value = lambda component: dir_loss(x_test, component, sol_test[component])

    LMS = LossMeanSquares
    o = data.options
    losses: List[LossMeanSquares] = []
    if o.use_collloss:
        losses += [LMS('PDE_MASS', lambda: PDE_MASS(), weight=1e1),
                   LMS('PDE_MOMU', lambda: PDE_MOM(0), weight=1e0),
                   LMS('PDE_MOMV', lambda: PDE_MOM(1), weight=1e0)]
    if o.use_boundary:
        losses += [LMS('BCD_u_x0', lambda: BC_D("SX", 0)), LMS('BCD_v_x0', lambda: BC_D("SX", 1)),
                   LMS('BCD_u_x1', lambda: BC_D("DX", 0)), LMS('BCD_v_x1', lambda: BC_D("DX", 1)),
                   LMS('BCD_u_y0', lambda: BC_D("BOT", 0)), LMS('BCD_v_y0', lambda: BC_D("BOT", 1)),
                   LMS('BCD_u_y1', lambda: BC_D("TOP", 0)), LMS('BCD_v_y1', lambda: BC_D("TOP", 1))]
    if o.use_initialc:  # cavity_unsteady.py:56,:244
        losses += [LMS('IC_u', lambda: IN_C(0)), LMS('IC_v', lambda: IN_C(1)), LMS('IC_p', lambda: IN_C(2))]
    losses += [LMS('Fit_u', lambda: fit_velocity(0)), LMS('Fit_v', lambda: fit_velocity(1))]   # Q3
    losses += [LMS('Fit_p', lambda: fit_pressure())]
    loss_test = [LMS('u_test', lambda: exact_value(0)), LMS('v_test', lambda: exact_value(1)),
                 LMS('p_test', lambda: exact_value(2))]
    return OptimizationProblem(model.variables, losses, loss_test)


# --------------------------------------------------------------------------------------------
# Colliding_Flow and Poiseuille_Flow share the operator-style closures
# --------------------------------------------------------------------------------------------

def _operator_style(data, variables, rho: float, mu: float, neumann_outflow: bool, with_fit_p: bool,
                    in_tape_divergence: bool = False) -> OptimizationProblem:
    model = KerasMLP(variables)
    dim = 2
    norm_vel, norm_pre = data.norm_vel, data.norm_pre
    x_pde = _t(data.x_pde)
    bnd_pts = {k: _t(v) for k, v in data.bnd_pts.items()}
    bnd_val = [{k: _t(v) for k, v in d.items()} for d in data.bnd_val]
    sol_noise = [_t(v) for v in data.sol_noise]
    sol_test = [_t(v) for v in data.sol_test]
    x_vel, x_pres, x_test = _t(data.x_vel), _t(data.x_pres), _t(data.x_test)
    gradient, divergence, laplacian = gradient_scalar, divergence_vector, laplacian_scalar

    def PDE_MASS():  # colliding_flow.py:160-165 / poiseuille_flow.py:173-178
        x = _watch(x_pde)
        with GradientTape(persistent=True) as tape:
            tape.watch(x)
            u_vect = model(x)[:, 0:2]
            if in_tape_divergence:  # the corrected form (Examples_Old/Poiseuille/poiseuille.py:98-103)
                return divergence(tape, u_vect, x, dim)
        return divergence(tape, u_vect, x, dim)   # called after the tape closed: Q1 -> zeros

    def PDE_MOM(k):  # colliding_flow.py:167-184 / poiseuille_flow.py:180-197
        x = _watch(x_pde)
        with GradientTape(persistent=True) as tape:
            tape.watch(x)
            u_vect = model(x)
            p = u_vect[:, 2] * norm_pre
            u_eq = u_vect[:, k] * norm_vel
            grad_eq = gradient(tape, u_eq, x)
            dp = gradient(tape, p, x)[:, k]
            deqx = grad_eq[:, 0]
            deqy = grad_eq[:, 1]
            lapl_eq = laplacian(tape, u_eq, x, dim)
            # un-scaled convecting velocity: Q4 (colliding_flow.py:181, poiseuille_flow.py:194)
            unnormed_lhs = rho * (u_vect[:, 0] * deqx + u_vect[:, 1] * deqy) - mu * (lapl_eq) + dp
            norm_const = 1 / max(norm_pre, norm_vel)
        return unnormed_lhs * norm_const

    def neu_loss(x, k, j, rhs=0):  # poiseuille_flow.py:199-209
        x = _watch(x)
        with GradientTape(persistent=True) as tape:
            tape.watch(x)
            uk = model(x)[:, k] * norm_vel
            p = model(x)[:, 2] * norm_pre
            uk_j = gradient(tape, uk, x)[:, j]
            norm_const = 1 / max(norm_pre, norm_vel)
        return norm_const * (uk_j * mu - p * (j == k) - rhs)

    def dir_loss(points, component, rhs):
        uk = model(points)[:, component]
        return uk - rhs

    BC_D = lambda edge, component: dir_loss(bnd_pts[edge], component, bnd_val[component][edge])
    BC_N = lambda edge, component, direction: neu_loss(bnd_pts[edge], component, direction, bnd_val[component][edge])
    fit_velocity = lambda component: dir_loss(x_vel, component, sol_noise[component])
    fit_pressure = lambda: dir_loss(x_pres, 2, sol_noise[2])
    exact_value = lambda component: dir_loss(x_test, component, sol_test[component])

    LMS = LossMeanSquares
    o = data.options
    losses: List[LossMeanSquares] = []
    if o.use_collloss:
        losses += [LMS('PDE_MASS', lambda: PDE_MASS(), weight=1e1),
                   LMS('PDE_MOMU', lambda: PDE_MOM(0), weight=1e0),
                   LMS('PDE_MOMV', lambda: PDE_MOM(1), weight=1e0)]
    if o.use_boundary:
        losses += [LMS('BCD_u_x0', lambda: BC_D("SX", 0)), LMS('BCD_v_x0', lambda: BC_D("SX", 1)),
                   LMS('BCD_u_y0', lambda: BC_D("BOT", 0)), LMS('BCD_v_y0', lambda: BC_D("BOT", 1)),
                   LMS('BCD_u_y1', lambda: BC_D("TOP", 0)), LMS('BCD_v_y1', lambda: BC_D("TOP", 1))]
        if neumann_outflow:  # poiseuille_flow.py:244-245,250
            losses += [LMS('BCN_u_x1', lambda: BC_N("DX", 0, 0)), LMS('BCN_v_x1', lambda: BC_N("DX", 1, 0))]
        else:                # colliding_flow.py:218-219
            losses += [LMS('BCD_u_x1', lambda: BC_D("DX", 0)), LMS('BCD_v_x1', lambda: BC_D("DX", 1))]
    losses += [LMS('Fit_u', lambda: fit_velocity(0)), LMS('Fit_v', lambda: fit_velocity(1))]   # Q3
    if with_fit_p:   # commented out in poiseuille_flow.py:254
        losses += [LMS('Fit_p', lambda: fit_pressure())]
    loss_test = [LMS('u_test', lambda: exact_value(0)), LMS('v_test', lambda: exact_value(1)),
                 LMS('p_test', lambda: exact_value(2))]
    return OptimizationProblem(model.variables, losses, loss_test)


def colliding_flow(data, variables, in_tape_divergence: bool = False) -> OptimizationProblem:
    return _operator_style(data, variables, rho=1.0, mu=1.0, neumann_outflow=False, with_fit_p=True,
                           in_tape_divergence=in_tape_divergence)


def poiseuille_flow(data, variables, in_tape_divergence: bool = False) -> OptimizationProblem:
    return _operator_style(data, variables, rho=data.consts["rho"], mu=data.consts["mu"],
                           neumann_outflow=True, with_fit_p=False, in_tape_divergence=in_tape_divergence)


# --------------------------------------------------------------------------------------------
# Colliding_Flow, pressure-mean variant  (Examples/Colliding_Flow/colliding_flow_pressmean.py)
# --------------------------------------------------------------------------------------------

def colliding_flow_pressmean(data, variables) -> OptimizationProblem:
    model = KerasMLP(variables)
    dim = 2
    vel_max, p_max = data.norm_vel, data.norm_pre
    x_PDE, x_col, x_BCD = _t(data.x_pde), _t(data.x_vel), _t(data.extra["x_BCD"])
    x_test, x_pres = _t(data.x_test), _t(data.x_pres)
    rhs = {k: _t(data.extra[k]) for k in ("bcd_u", "bcd_v", "col_u", "col_v", "col_p")}
    sol_test = [_t(v) for v in data.sol_test]

    def PDE_MASS(x):  # colliding_flow_pressmean.py:140-145
        x = _watch(x)
        with GradientTape(persistent=True) as tape:
            tape.watch(x)
            u_vect = model(x)[:, 0:2] * vel_max
            div = divergence_vector(tape, u_vect, x, dim)
        return div

    def PDE_MOM(x, k):  # :147-159 (zero forcing)
        x = _watch(x)
        with GradientTape(persistent=True) as tape:
            tape.watch(x)
            u_vect = model(x)
            p = u_vect[:, 2] * p_max
            u_eq = u_vect[:, k] * vel_max
            dp = gradient_scalar(tape, p, x)[:, k]
            lapl_eq = laplacian_scalar(tape, u_eq, x, dim)
        return - (lapl_eq) + dp

    def BC_D(x, k, target):  # :163-166 and exact_value :170-173; target = (g_bc(x) + noise) / norm
        return model(x)[:, k] - target

    def PRESS_0(x):  # :176-179
        uk = model(x)[:, 2]
        return torch.abs(torch.mean(uk))

    LMS = LossMeanSquares
    losses = [LMS('PDE_MASS', lambda: PDE_MASS(x_PDE), normalization=1e4, weight=1e0),
              LMS('PDE_MOMU', lambda: PDE_MOM(x_PDE, 0), normalization=1e4, weight=1e-2),
              LMS('PDE_MOMV', lambda: PDE_MOM(x_PDE, 1), normalization=1e4, weight=1e-2),
              LMS('BCD_u', lambda: BC_D(x_BCD, 0, rhs["bcd_u"]), weight=1e0),
              LMS('BCD_v', lambda: BC_D(x_BCD, 1, rhs["bcd_v"]), weight=1e0)]
    if data.consts["collocation"]:
        losses += [LMS('COL_u', lambda: BC_D(x_col, 0, rhs["col_u"])), LMS('COL_v', lambda: BC_D(x_col, 1, rhs["col_v"]))]
    if data.consts["press_mode"] == "Collocation":
        losses += [LMS('COL_p', lambda: BC_D(x_pres, 2, rhs["col_p"]))]
    if data.consts["press_mode"] == "Mean":
        losses += [Loss('PRESS_0', lambda: PRESS_0(x_pres), normalization=1e0, weight=1e-2, non_negative=True)]
    loss_test = [LMS('u_fit', lambda: BC_D(x_test, 0, sol_test[0])), LMS('v_fit', lambda: BC_D(x_test, 1, sol_test[1])),
                 LMS('p_fit', lambda: BC_D(x_test, 2, sol_test[2]))]
    return OptimizationProblem(model.variables, losses, loss_test)


# --------------------------------------------------------------------------------------------
# Coronary_Flow  (Examples/Coronary_Flow/coronary_flow_steady.py)
# --------------------------------------------------------------------------------------------

def coronary_flow(data, variables, in_tape_neumann: bool = False) -> OptimizationProblem:
    model = KerasMLP(variables)
    norm_vel, norm_pre, ni = data.norm_vel, data.norm_pre, data.consts["ni"]
    x_pde = _t(data.x_pde)
    bnd_pts = {k: _t(v) for k, v in data.bnd_pts.items()}
    bnd_val = [{k: _t(v) for k, v in d.items()} for d in data.bnd_val]
    sol_noise = [_t(v) for v in data.sol_noise]
    sol_test = [_t(v) for v in data.sol_test]
    x_vel, x_test = _t(data.x_vel), _t(data.x_test)
    gradient = gradient_scalar

    def PDE_MASS():  # coronary_flow_steady.py:163-170
        x = _watch(x_pde)
        with GradientTape(persistent=True) as tape:
            tape.watch(x)
            u_vect = model(x)[:, 0:2]
            du_x = gradient(tape, u_vect[:, 0], x)[:, 0]
            dv_y = gradient(tape, u_vect[:, 1], x)[:, 1]
        return du_x + dv_y

    def PDE_MOM(k):  # :172-192
        x = _watch(x_pde)
        with GradientTape(persistent=True) as tape:
            tape.watch(x)
            u_vect = model(x)
            p = u_vect[:, 2] * norm_pre
            u_eq = u_vect[:, k] * norm_vel
            dp = gradient(tape, p, x)[:, k]
            du_x = gradient(tape, u_eq, x)[:, 0]
            du_y = gradient(tape, u_eq, x)[:, 1]
            du_xx = gradient(tape, du_x, x)[:, 0]
            du_yy = gradient(tape, du_y, x)[:, 1]
            conv1 = torch.mul(norm_vel * u_vect[:, 0], du_x)
            conv2 = torch.mul(norm_vel * u_vect[:, 1], du_y)
            unnormed_lhs = - ni * (du_xx + du_yy) + dp + conv1 + conv2
            norm_const = 1 / max(norm_pre, norm_vel)
        return unnormed_lhs * norm_const

    def dir_loss(points, component, rhs):  # :196-199
        uk = model(points)[:, component]
        return uk - rhs

    def neu_loss(edge, k, rhs):  # :201-215
        x = _watch(bnd_pts[edge])
        n = torch.tensor([2.0, 1.0] if edge == 'OUT1' else [1.0, 0.0], dtype=DTYPE).reshape(2, 1)
        with GradientTape(persistent=True) as tape:
            tape.watch(x)
            if in_tape_neumann:      # the intended traction condition
                u_vect = model(x)
        if not in_tape_neumann:
            # the script calls the model AFTER the tape closed (:209): TF records nothing, so the gradient below
            # is unconnected.  torch has no tape, so the same is stated by cutting x out of the graph.
            u_vect = model(x.detach())
        u_eq = u_vect[:, k] * norm_vel
        p_eq = u_vect[:, 2] * norm_pre
        grad = gradient(tape, u_eq, x)[:, 0:2]
        # the script subtracts an [N] vector from the [N,1] product (:215), which broadcasts to [N,N]; with the zero
        # gradient every row of that matrix is the same vector, so its mean square is the vector's.
        return ni * torch.matmul(grad, n).reshape(-1) - p_eq * n[k] - rhs

    BC_D = lambda edge, component: dir_loss(bnd_pts[edge], component, bnd_val[component][edge])
    BC_N = lambda edge, component: neu_loss(edge, component, bnd_val[component][edge])
    fit_velocity = lambda component: dir_loss(x_vel, component, sol_noise[component])
    exact_value = lambda component: dir_loss(x_test, component, sol_test[component])

    LMS = LossMeanSquares
    o = data.options
    losses: List[LossMeanSquares] = []
    if o.use_collloss:
        losses += [LMS('PDE_MASS', lambda: PDE_MASS(), weight=1e2),
                   LMS('PDE_MOMU', lambda: PDE_MOM(0), weight=1e1),
                   LMS('PDE_MOMV', lambda: PDE_MOM(1), weight=1e1)]
    if o.use_boundary:
        losses += [LMS('BCD_u_NS', lambda: BC_D("NOSL", 0)), LMS('BCD_v_NS', lambda: BC_D("NOSL", 1)),
                   LMS('BCD_u_IN', lambda: BC_D("INF", 0)), LMS('BCD_v_IN', lambda: BC_D("INF", 1)),
                   LMS('BCN_u_OUT1', lambda: BC_N("OUT1", 0), weight=1e-3),
                   LMS('BCN_v_OUT1', lambda: BC_N("OUT1", 1), weight=1e-3),
                   LMS('BCN_u_OUT2', lambda: BC_N("OUT2", 0), weight=1e-3),
                   LMS('BCN_v_OUT2', lambda: BC_N("OUT2", 1), weight=1e-3)]
    losses += [LMS('Fit_u', lambda: fit_velocity(0)), LMS('Fit_v', lambda: fit_velocity(1))]   # Q3; Fit_p commented out
    loss_test = [LMS('u_test', lambda: exact_value(0)), LMS('v_test', lambda: exact_value(1)),
                 LMS('p_test', lambda: exact_value(2))]
    return OptimizationProblem(model.variables, losses, loss_test)


# --------------------------------------------------------------------------------------------
# Poisson  (Examples/Poisson_Problem/poisson.py, poisson_misto.py)
# --------------------------------------------------------------------------------------------

def poisson(data, variables) -> OptimizationProblem:
    model = KerasMLP(variables)
    dim = 2
    x_PDE = _t(data.x_pde)
    f = _t(data.extra["f"])
    x_test, u_test = _t(data.x_test), _t(data.extra["u_test"])[:, None]

    def PDE():  # poisson.py:58-63
        x = _watch(x_PDE)
        with GradientTape(persistent=True) as tape:
            tape.watch(x)
            u = model(x)
            laplacian = laplacian_scalar(tape, u, x, dim)
        return -laplacian - f

    if data.name == "poisson":
        x_BC = _t(data.extra["x_BC"])
        losses = [LossMeanSquares('PDE', PDE, weight=2.0),
                  LossMeanSquares('BC', lambda: model(x_BC))]
    else:
        x_BC_D, x_BC_N, g = _t(data.extra["x_BC_D"]), _t(data.extra["x_BC_N"]), _t(data.extra["g"])

        def BC_D():  # poisson_misto.py:69-73
            return model(x_BC_D)

        def BC_N():  # poisson_misto.py:75-80
            x = _watch(x_BC_N)
            with GradientTape(persistent=True) as tape:
                tape.watch(x)
                u = model(x)
                u_x = gradient_scalar(tape, u, x)[:, 0]
            return u_x - g

        losses = [LossMeanSquares('PDE', PDE, weight=1e2),
                  LossMeanSquares('BC_D', BC_D),
                  LossMeanSquares('BC_N', BC_N)]
    loss_test = LossMeanSquares('fit', lambda: model(x_test) - u_test)
    return OptimizationProblem(model.variables, losses, loss_test)


BUILDERS = {
    "cavity_steady": cavity_steady,
    "cavity_unsteady": cavity_unsteady,
    "colliding_flow": colliding_flow,
    "poiseuille_flow": poiseuille_flow,
    "poisson": poisson,
    "poisson_misto": poisson,
    "coronary_flow": coronary_flow,
    "colliding_flow_pressmean": colliding_flow_pressmean,
}


def build(data, variables, **kw) -> OptimizationProblem:
    return BUILDERS[data.name](data, variables, **kw)


def glorot_uniform_variables(dim: int, hidden: Sequence[int], out_dim: int, seed: int = 1,
                             bias_std: float = 0.0) -> List[torch.Tensor]:
    """GlorotUniform kernels / zero biases (Model.json), optional N(0, bias_std^2) biases for the
    "mid-training" parity variant.  Values are rounded to float32."""
    rng = np.random.default_rng(seed)
    sizes = [dim] + list(hidden) + [out_dim]
    out = []
    for i in range(len(sizes) - 1):
        lim = np.sqrt(6.0 / (sizes[i] + sizes[i + 1]))
        K = rng.uniform(-lim, lim, size=(sizes[i], sizes[i + 1]))
        b = rng.standard_normal(sizes[i + 1]) * bias_std
        out += [torch.as_tensor(K.astype(np.float32).astype(np.float64)),
                torch.as_tensor(b.astype(np.float32).astype(np.float64))]
    return out
