"""The loss tables of the in-scope example scripts (the five BASELINE configs and Coronary_Flow), written against the facade.

Each function is the "Loss Building" + "Model's Setup" section of one script with the closures
replaced by their declarative forms; names, weights and order are the script's.  ``faithful=True``
keeps the reference's quirks: the identically-zero mass residual of
Colliding/Poiseuille (Q1), Cavity_Steady's viscous sign (Q2), fit terms always present (Q3), the
un-scaled convecting velocity (Q4).  ``faithful=False`` gives the mathematically intended residuals
(in-tape divergence, -laplacian); the benchmark uses the in-tape mass residual so the full
Navier-Stokes residual is exercised.
"""
from __future__ import annotations

from typing import List, Tuple

from . import residuals as R
from .api import Loss
from .api import LossMeanSquares as LMS
from .problems import ProblemData
from .residuals import PointSet


def _common_sets(data: ProblemData):
    sets = {"PDE": PointSet(data.x_pde, "PDE")}
    for edge, pts in data.bnd_pts.items():
        sets[edge] = PointSet(pts, edge)
    sets["Vel"] = PointSet(data.x_vel, "Vel")
    sets["Pres"] = PointSet(data.x_pres, "Pres")
    sets["Test"] = PointSet(data.x_test, "Test")
    return sets


def _dirichlet_edges(data, sets, order) -> List[LMS]:
    tag = {"SX": "x0", "DX": "x1", "BOT": "y0", "TOP": "y1"}
    out = []
    for edge in order:
        for comp, cname in ((0, "u"), (1, "v")):
            out.append(LMS(f"BCD_{cname}_{tag[edge]}",
                           lambda e=edge, c=comp: R.dirichlet(sets[e], c, data.bnd_val[c][e]), weight=1e0))
    return out


def _fit_and_test(data, sets, with_fit_p=True) -> Tuple[List[LMS], List[LMS]]:
    fit = [LMS("Fit_u", lambda: R.dirichlet(sets["Vel"], 0, data.sol_noise[0]), weight=1e0),
           LMS("Fit_v", lambda: R.dirichlet(sets["Vel"], 1, data.sol_noise[1]), weight=1e0)]
    if with_fit_p:
        fit.append(LMS("Fit_p", lambda: R.dirichlet(sets["Pres"], 2, data.sol_noise[2]), weight=1e0))
    test = [LMS(f"{c}_test", lambda i=i: R.dirichlet(sets["Test"], i, data.sol_test[i]))
            for i, c in enumerate(("u", "v", "p"))]
    return fit, test


def cavity_steady(data: ProblemData, faithful: bool = True):
    """cavity_steady.py:155-235"""
    s = _common_sets(data)
    nv, npr, o = data.norm_vel, data.norm_pre, data.options
    vxx = +1.0 if faithful else -1.0          # Q2: `du_xx - du_yy` (:185)
    mom = lambda k: R.momentum(s["PDE"], k, nv, npr, conv_scale=nv, visc_xx=vxx, visc_yy=-1.0)
    losses: List[LMS] = []
    if o.use_collloss:
        losses += [LMS("PDE_MASS", lambda: R.mass(s["PDE"]), weight=1e1),
                   LMS("PDE_MOMU", lambda: mom(0), weight=1e0),
                   LMS("PDE_MOMV", lambda: mom(1), weight=1e0)]
    if o.use_boundary:
        losses += _dirichlet_edges(data, s, ("SX", "DX", "BOT", "TOP"))
    fit, test = _fit_and_test(data, s)
    if faithful:
        losses += fit                           # Q3: always added (:198-199 shadow the flags)
    else:
        losses += (fit[:2] if o.fit_velocity else []) + (fit[2:] if o.fit_pressure else [])
    return losses, test


def cavity_unsteady(data: ProblemData, faithful: bool = True):
    """cavity_unsteady.py:165-250"""
    s = _common_sets(data)
    nv, npr, o = data.norm_vel, data.norm_pre, data.options
    mom = lambda k: R.momentum(s["PDE"], k, nv, npr, conv_scale=nv, visc_xx=-1.0, visc_yy=-1.0,
                               time_derivative=True)
    losses: List[LMS] = []
    if o.use_collloss:
        losses += [LMS("PDE_MASS", lambda: R.mass(s["PDE"]), weight=1e1),
                   LMS("PDE_MOMU", lambda: mom(0), weight=1e0),
                   LMS("PDE_MOMV", lambda: mom(1), weight=1e0)]
    if o.use_boundary:
        losses += _dirichlet_edges(data, s, ("SX", "DX", "BOT", "TOP"))
    if o.use_initialc:
        losses += [LMS(f"IC_{c}", lambda i=i: R.dirichlet(s["IC"], i, None), weight=1e0)
                   for i, c in enumerate(("u", "v", "p"))]
    fit, test = _fit_and_test(data, s)
    if faithful:
        losses += fit
    else:
        losses += (fit[:2] if o.fit_velocity else []) + (fit[2:] if o.fit_pressure else [])
    return losses, test


def colliding_flow(data: ProblemData, faithful: bool = True):
    """colliding_flow.py:152-232"""
    s = _common_sets(data)
    nv, npr, o = data.norm_vel, data.norm_pre, data.options
    mom = lambda k: R.momentum(s["PDE"], k, nv, npr, conv_scale=1.0 if faithful else nv,
                               visc_xx=-1.0, visc_yy=-1.0)
    losses: List[LMS] = []
    if o.use_collloss:
        losses += [LMS("PDE_MASS", lambda: R.mass(s["PDE"], in_tape=not faithful), weight=1e1),
                   LMS("PDE_MOMU", lambda: mom(0), weight=1e0),
                   LMS("PDE_MOMV", lambda: mom(1), weight=1e0)]
    if o.use_boundary:
        losses += _dirichlet_edges(data, s, ("SX", "BOT", "TOP", "DX"))
    fit, test = _fit_and_test(data, s)
    losses += fit if faithful else ((fit[:2] if o.fit_velocity else []) + (fit[2:] if o.fit_pressure else []))
    return losses, test


def poiseuille_flow(data: ProblemData, faithful: bool = True):
    """poiseuille_flow.py:165-258"""
    s = _common_sets(data)
    nv, npr, o = data.norm_vel, data.norm_pre, data.options
    rho, mu = data.consts["rho"], data.consts["mu"]
    mom = lambda k: R.momentum(s["PDE"], k, nv, npr, conv_scale=rho if faithful else rho * nv,
                               visc_xx=-mu, visc_yy=-mu)
    losses: List[LMS] = []
    if o.use_collloss:
        losses += [LMS("PDE_MASS", lambda: R.mass(s["PDE"], in_tape=not faithful), weight=1e1),
                   LMS("PDE_MOMU", lambda: mom(0), weight=1e0),
                   LMS("PDE_MOMV", lambda: mom(1), weight=1e0)]
    if o.use_boundary:
        losses += _dirichlet_edges(data, s, ("SX", "BOT", "TOP"))
        losses += [LMS("BCN_u_x1", lambda: R.neumann(s["DX"], 0, 0, data.bnd_val[0]["DX"], nv, npr, mu), weight=1e0),
                   LMS("BCN_v_x1", lambda: R.neumann(s["DX"], 1, 0, data.bnd_val[1]["DX"], nv, npr, mu), weight=1e0)]
    fit, test = _fit_and_test(data, s, with_fit_p=False)   # Fit_p commented out (:254)
    losses += fit if faithful else (fit if o.fit_velocity else [])
    return losses, test


def coronary_flow(data: ProblemData, faithful: bool = True):
    """coronary_flow_steady.py:159-246"""
    s = _common_sets(data)
    nv, npr, o, ni = data.norm_vel, data.norm_pre, data.options, data.consts["ni"]
    mom = lambda k: R.momentum(s["PDE"], k, nv, npr, conv_scale=nv, visc_xx=-ni, visc_yy=-ni)
    normal = {"OUT1": (2.0, 1.0), "OUT2": (1.0, 0.0)}     # :199-204
    bcn = lambda e, k: R.outflow_stress(s[e], k, normal[e], data.bnd_val[k][e], nv, npr, ni, in_tape=not faithful)
    losses: List[LMS] = []
    if o.use_collloss:
        losses += [LMS("PDE_MASS", lambda: R.mass(s["PDE"]), weight=1e2),
                   LMS("PDE_MOMU", lambda: mom(0), weight=1e1),
                   LMS("PDE_MOMV", lambda: mom(1), weight=1e1)]
    if o.use_boundary:
        losses += [LMS("BCD_u_NS", lambda: R.dirichlet(s["NOSL"], 0, data.bnd_val[0]["NOSL"]), weight=1e0),
                   LMS("BCD_v_NS", lambda: R.dirichlet(s["NOSL"], 1, data.bnd_val[1]["NOSL"]), weight=1e0),
                   LMS("BCD_u_IN", lambda: R.dirichlet(s["INF"], 0, data.bnd_val[0]["INF"]), weight=1e0),
                   LMS("BCD_v_IN", lambda: R.dirichlet(s["INF"], 1, data.bnd_val[1]["INF"]), weight=1e0),
                   LMS("BCN_u_OUT1", lambda: bcn("OUT1", 0), weight=1e-3),
                   LMS("BCN_v_OUT1", lambda: bcn("OUT1", 1), weight=1e-3),
                   LMS("BCN_u_OUT2", lambda: bcn("OUT2", 0), weight=1e-3),
                   LMS("BCN_v_OUT2", lambda: bcn("OUT2", 1), weight=1e-3)]
    fit, test = _fit_and_test(data, s, with_fit_p=False)    # Fit_p commented out (:240)
    losses += fit if faithful else (fit if o.fit_velocity else [])
    return losses, test


def colliding_flow_pressmean(data: ProblemData, faithful: bool = True):
    """colliding_flow_pressmean.py:137-214"""
    vel_max, p_max = data.norm_vel, data.norm_pre
    pde, bcd = PointSet(data.x_pde, "PDE"), PointSet(data.extra["x_BCD"], "BCD")
    col, pres, test = PointSet(data.x_vel, "col"), PointSet(data.x_pres, "pres"), PointSet(data.x_test, "Test")
    losses = [LMS("PDE_MASS", lambda: R.mass(pde, scale=vel_max), normalization=1e4, weight=1e0),
              LMS("PDE_MOMU", lambda: R.stokes_momentum(pde, 0, vel_max, p_max), normalization=1e4, weight=1e-2),
              LMS("PDE_MOMV", lambda: R.stokes_momentum(pde, 1, vel_max, p_max), normalization=1e4, weight=1e-2),
              LMS("BCD_u", lambda: R.dirichlet(bcd, 0, data.extra["bcd_u"]), weight=1e0),
              LMS("BCD_v", lambda: R.dirichlet(bcd, 1, data.extra["bcd_v"]), weight=1e0)]
    if data.consts["collocation"]:
        losses += [LMS("COL_u", lambda: R.dirichlet(col, 0, data.extra["col_u"]), weight=1e0),
                   LMS("COL_v", lambda: R.dirichlet(col, 1, data.extra["col_v"]), weight=1e0)]
    if data.consts["press_mode"] == "Collocation":
        losses += [LMS("COL_p", lambda: R.dirichlet(pres, 2, data.extra["col_p"]), weight=1e0)]
    if data.consts["press_mode"] == "Mean":
        losses += [Loss("PRESS_0", lambda: R.mean_value(pres, 2), normalization=1e0, weight=1e-2, non_negative=True)]
    loss_test = [LMS(f"{c}_fit", lambda i=i: R.dirichlet(test, i, data.sol_test[i])) for i, c in enumerate(("u", "v", "p"))]
    return losses, loss_test


def poisson(data: ProblemData, faithful: bool = True):
    """poisson.py:58-69 / poisson_misto.py:62-88"""
    pde = PointSet(data.x_pde, "PDE")
    test_set = PointSet(data.x_test, "Test")
    if data.name == "poisson":
        bc = PointSet(data.extra["x_BC"], "BC")
        losses = [LMS("PDE", lambda: R.poisson_pde(pde, data.extra["f"]), weight=2.0),
                  LMS("BC", lambda: R.dirichlet(bc, 0, None))]
    else:
        bcd, bcn = PointSet(data.extra["x_BC_D"], "BC_D"), PointSet(data.extra["x_BC_N"], "BC_N")
        losses = [LMS("PDE", lambda: R.poisson_pde(pde, data.extra["f"]), weight=1e2),
                  LMS("BC_D", lambda: R.dirichlet(bcd, 0, None)),
                  LMS("BC_N", lambda: R.normal_derivative(bcn, 0, 0, data.extra["g"]))]
    loss_test = LMS("fit", lambda: R.dirichlet(test_set, 0, data.extra["u_test"]))
    return losses, [loss_test]


TABLES = {
    "cavity_steady": cavity_steady,
    "cavity_unsteady": cavity_unsteady,
    "colliding_flow": colliding_flow,
    "poiseuille_flow": poiseuille_flow,
    "coronary_flow": coronary_flow,
    "colliding_flow_pressmean": colliding_flow_pressmean,
    "poisson": poisson,
    "poisson_misto": poisson,
}


def build_loss_table(data: ProblemData, faithful: bool = True):
    return TABLES[data.name](data, faithful=faithful)
