"""Strong-Wolfe line search of the BFGS round (``ns.minimize(pb, 'scipy', 'BFGS', ...)``, cavity_steady.py:247).

nisaba hands the round to ``scipy.optimize.minimize(method='BFGS')``, whose step lengths come from the More-Thuente
search of MINPACK-2 (``dcsrch`` / ``dcstep``; J. J. More, D. J. Thuente, "Line search algorithms with guaranteed
sufficient decrease", ACM TOMS 20 (1994) 286-307) and, when that gives up, from the bracketing / zoom search of
Nocedal & Wright (Alg. 3.5 / 3.6), which SciPy publishes as ``scipy.optimize.line_search``.  This module restates the
More-Thuente algorithm from the paper so that the round needs no private SciPy symbol; it works on the scalar function
``phi(alpha) -> (value, slope)`` -- one device loss step delivers both -- and the iterates it produces are checked against
``scipy.optimize.minimize`` itself in tests/test_bfgs.py.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np


def _mt_step(stx, fx, dx, sty, fy, dy, stp, fp, dp, brackt, stpmin, stpmax):
    """One safeguarded step of More-Thuente (section 4 of the paper): updates the interval of uncertainty
    [stx, sty] and returns the next trial step."""
    sgnd = dp * (dx / abs(dx))
    if fp > fx:
        # case 1: a higher function value -- the minimum is bracketed; cubic unless the quadratic step is closer to stx
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp
        s = max(abs(theta), abs(dx), abs(dp))
        gamma = s * np.sqrt((theta / s) ** 2 - (dx / s) * (dp / s))
        if stp < stx:
            gamma = -gamma
        p = (gamma - dx) + theta
        q = ((gamma - dx) + gamma) + dp
        r = p / q
        stpc = stx + r * (stp - stx)
        stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx)
        stpf = stpc if abs(stpc - stx) < abs(stpq - stx) else stpc + (stpq - stpc) / 2.0
        brackt = True
    elif sgnd < 0.0:
        # case 2: lower value, derivatives of opposite sign -- bracketed; the step farther from stp of cubic / secant
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp
        s = max(abs(theta), abs(dx), abs(dp))
        gamma = s * np.sqrt((theta / s) ** 2 - (dx / s) * (dp / s))
        if stp > stx:
            gamma = -gamma
        p = (gamma - dp) + theta
        q = ((gamma - dp) + gamma) + dx
        r = p / q
        stpc = stp + r * (stx - stp)
        stpq = stp + (dp / (dp - dx)) * (stx - stp)
        stpf = stpc if abs(stpc - stp) > abs(stpq - stp) else stpq
        brackt = True
    elif abs(dp) < abs(dx):
        # case 3: lower value, same sign, the derivative shrinks -- the cubic may have no minimiser in the right direction
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp
        s = max(abs(theta), abs(dx), abs(dp))
        gamma = s * np.sqrt(max(0.0, (theta / s) ** 2 - (dx / s) * (dp / s)))
        if stp > stx:
            gamma = -gamma
        p = (gamma - dp) + theta
        q = (gamma + (dx - dp)) + gamma
        r = p / q
        if r < 0.0 and gamma != 0.0:
            stpc = stp + r * (stx - stp)
        elif stp > stx:
            stpc = stpmax
        else:
            stpc = stpmin
        stpq = stp + (dp / (dp - dx)) * (stx - stp)
        if brackt:
            stpf = stpc if abs(stpc - stp) < abs(stpq - stp) else stpq
            if stp > stx:
                stpf = min(stp + 0.66 * (sty - stp), stpf)
            else:
                stpf = max(stp + 0.66 * (sty - stp), stpf)
        else:
            stpf = stpc if abs(stpc - stp) > abs(stpq - stp) else stpq
            stpf = max(stpmin, min(stpmax, stpf))
    else:
        # case 4: lower value, same sign, the derivative does not shrink
        if brackt:
            theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp
            s = max(abs(theta), abs(dy), abs(dp))
            gamma = s * np.sqrt((theta / s) ** 2 - (dy / s) * (dp / s))
            if stp > sty:
                gamma = -gamma
            p = (gamma - dp) + theta
            q = ((gamma - dp) + gamma) + dy
            r = p / q
            stpf = stp + r * (sty - stp)
        elif stp > stx:
            stpf = stpmax
        else:
            stpf = stpmin
    if fp > fx:
        sty, fy, dy = stp, fp, dp
    else:
        if sgnd < 0.0:
            sty, fy, dy = stx, fx, dx
        stx, fx, dx = stp, fp, dp
    return stx, fx, dx, sty, fy, dy, stpf, brackt


def more_thuente(phi: Callable[[float], Tuple[float, float]], phi0: float, derphi0: float, alpha1: float, ftol: float = 1e-4,
                 gtol: float = 0.9, xtol: float = 1e-14, stpmin: float = 1e-100, stpmax: float = 1e100,
                 maxiter: int = 100) -> Tuple[Optional[float], float, float]:
    """Step ``alpha`` with phi(alpha) <= phi0 + ftol alpha phi'(0) and |phi'(alpha)| <= gtol |phi'(0)|.
    Returns (alpha or None when the search stops on a warning / error, phi(alpha), phi'(alpha))."""
    if not (alpha1 >= stpmin and alpha1 <= stpmax) or derphi0 >= 0.0:
        return None, phi0, derphi0
    xtrapl, xtrapu = 1.1, 4.0
    brackt, stage = False, 1
    finit, ginit = phi0, derphi0
    gtest = ftol * ginit
    width = stpmax - stpmin
    width1 = width / 0.5
    stx, fx, gx = 0.0, finit, ginit
    sty, fy, gy = 0.0, finit, ginit
    stmin, stmax = 0.0, alpha1 + xtrapu * alpha1
    stp = alpha1
    f = g = 0.0
    for _ in range(maxiter):
        f, g = phi(stp)
        ftest = finit + stp * gtest
        if stage == 1 and f <= ftest and g >= 0.0:
            stage = 2
        # convergence wins over the warnings (rounding errors, xtol, step at a bound): those return no step
        if f <= ftest and abs(g) <= gtol * (-ginit):
            return stp, f, g
        if brackt and (stp <= stmin or stp >= stmax):
            return None, f, g
        if brackt and stmax - stmin <= xtol * stmax:
            return None, f, g
        if stp == stpmax and f <= ftest and g <= gtest:
            return None, f, g
        if stp == stpmin and (f > ftest or g >= gtest):
            return None, f, g
        if stage == 1 and f <= fx and f > ftest:
            # first stage: the modified function psi(a) = phi(a) - phi(0) - ftol a phi'(0) (section 3 of the paper)
            fm, fxm, fym = f - stp * gtest, fx - stx * gtest, fy - sty * gtest
            gm, gxm, gym = g - gtest, gx - gtest, gy - gtest
            stx, fxm, gxm, sty, fym, gym, stp, brackt = _mt_step(stx, fxm, gxm, sty, fym, gym, stp, fm, gm, brackt, stmin, stmax)
            fx, fy = fxm + stx * gtest, fym + sty * gtest
            gx, gy = gxm + gtest, gym + gtest
        else:
            stx, fx, gx, sty, fy, gy, stp, brackt = _mt_step(stx, fx, gx, sty, fy, gy, stp, f, g, brackt, stmin, stmax)
        if brackt:
            if abs(sty - stx) >= 0.66 * width1:
                stp = stx + 0.5 * (sty - stx)
            width1 = width
            width = abs(sty - stx)
            stmin, stmax = min(stx, sty), max(stx, sty)
        else:
            stmin = stp + xtrapl * (stp - stx)
            stmax = stp + xtrapu * (stp - stx)
        stp = min(max(stp, stpmin), stpmax)
        if (brackt and (stp <= stmin or stp >= stmax)) or (brackt and stmax - stmin <= xtol * stmax):
            stp = stx
    return None, f, g


def first_step(phi0: float, old_phi0: Optional[float], derphi0: float) -> float:
    """SciPy's initial trial step of a BFGS line search: the step that would repeat the previous decrease, at most 1."""
    if old_phi0 is not None and derphi0 != 0.0:
        a = min(1.0, 1.01 * 2.0 * (phi0 - old_phi0) / derphi0)
        return a if a >= 0.0 else 1.0
    return 1.0
