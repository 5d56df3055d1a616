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
