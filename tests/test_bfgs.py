"""The BFGS round: our driver with the O(P^2) inverse-Hessian update against
scipy.optimize.minimize(method='BFGS') -- the function nisaba's ``ns.minimize(pb, 'scipy', 'BFGS', n)`` calls
(cavity_steady.py:247) -- on problems whose f / gradient are exact float64 on the host."""
import numpy as np
import pytest
import scipy.optimize

from pinns_fluid_dynamics_b200.bfgs import minimize_bfgs


def _rosen(x):
    return scipy.optimize.rosen(x), scipy.optimize.rosen_der(x)


@pytest.mark.parametrize("n,iters", [(2, 12), (10, 25), (40, 30)])
def test_iterates_match_scipy_on_rosenbrock(n, iters):
    rng = np.random.default_rng(n)
    x0 = rng.uniform(-1.0, 1.0, n)
    ours = minimize_bfgs(_rosen, x0, maxiter=iters, device="cpu")
    ref = scipy.optimize.minimize(_rosen, x0, jac=True, method="BFGS", options={"maxiter": iters})
    assert ours.nit == ref.nit
    assert np.allclose(ours.x, ref.x, rtol=1e-7, atol=1e-9)
    assert abs(ours.fun - ref.fun) <= 1e-8 * max(1.0, abs(ref.fun))
    assert np.allclose(ours.hess_inv.numpy(), ref.hess_inv, rtol=1e-6, atol=1e-8)


def test_converges_and_reports_like_scipy():
    x0 = np.array([-1.2, 1.0, -0.5, 0.8])
    seen = []
    ours = minimize_bfgs(_rosen, x0, callback=lambda x: seen.append(x.copy()), device="cpu")
    ref = scipy.optimize.minimize(_rosen, x0, jac=True, method="BFGS")
    assert ours.success and ref.success
    assert ours.nit == ref.nit == len(seen)
    assert np.allclose(ours.x, np.ones(4), atol=1e-5)
    assert ours.nfev == ref.nfev            # one (f, g) evaluation per line-search trial point, cached like SciPy's


def test_quadratic_matches_scipy_including_inverse_hessian():
    rng = np.random.default_rng(0)
    A = rng.standard_normal((6, 6))
    A = A @ A.T + 6 * np.eye(6)
    b = rng.standard_normal(6)
    fun = lambda x: (0.5 * x @ A @ x - b @ x, A @ x - b)
    res = minimize_bfgs(fun, np.zeros(6), device="cpu", gtol=1e-10)
    ref = scipy.optimize.minimize(fun, np.zeros(6), jac=True, method="BFGS", options={"gtol": 1e-10})
    assert np.allclose(res.x, np.linalg.solve(A, b), atol=1e-8)
    assert res.nit == ref.nit
    assert np.allclose(res.hess_inv.numpy(), ref.hess_inv, rtol=1e-6, atol=1e-9)


# ---- the More-Thuente search on its own ---------------------------------------------------------------------------------

def _wolfe_ok(phi, a, f0, g0, c1=1e-4, c2=0.9):
    f, g = phi(a)
    return f <= f0 + c1 * a * g0 and abs(g) <= c2 * abs(g0)


@pytest.mark.parametrize("alpha1", [1e-3, 0.1, 1.0, 10.0, 1e3])
def test_more_thuente_returns_strong_wolfe_steps(alpha1):
    """functions 1-3 of the More-Thuente paper (table 1-3 there) from starting steps over six decades"""
    from pinns_fluid_dynamics_b200.linesearch import more_thuente
    b = 2.0
    f1 = lambda a: (-a / (a * a + b), (a * a - b) / (a * a + b) ** 2)
    b2 = 0.004
    f2 = lambda a: ((a + b2) ** 5 - 2 * (a + b2) ** 4, 5 * (a + b2) ** 4 - 8 * (a + b2) ** 3)
    for phi in (f1, f2):
        f0, g0 = phi(0.0)
        assert g0 < 0
        a, fa, ga = more_thuente(phi, f0, g0, alpha1, ftol=1e-4, gtol=0.9)
        assert a is not None and a > 0
        assert _wolfe_ok(phi, a, f0, g0)
        assert (fa, ga) == phi(a)


def test_more_thuente_rejects_ascent_directions_and_counts_evaluations():
    from pinns_fluid_dynamics_b200.linesearch import first_step, more_thuente
    calls = []

    def phi(a):
        calls.append(a)
        return (a - 1.0) ** 2, 2.0 * (a - 1.0)
    assert more_thuente(phi, 1.0, +2.0, 1.0)[0] is None and not calls          # slope >= 0: no step, no evaluation
    a, fa, ga = more_thuente(phi, 1.0, -2.0, 1.0)
    assert a == 1.0 and calls == [1.0]                                            # the first trial already satisfies both conditions
    # SciPy's first trial step of a BFGS line search: min(1, 1.01 * 2 (f - f_old) / slope), 1 when that is negative
    assert first_step(1.0, 1.5, -2.0) == pytest.approx(min(1.0, 1.01 * 2 * (1.0 - 1.5) / -2.0))
    assert first_step(1.0, 0.5, -2.0) == 1.0 and first_step(1.0, None, -2.0) == 1.0
