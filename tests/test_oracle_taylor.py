"""Taylor-mode jets + hand-written reverse sweep (oracle/taylor.py, the algorithm the kernels
implement) == nested reverse-mode autodiff (oracle/reference_step.py, the reference's formulation),
in float64, for every in-scope script."""
import os

import numpy as np
import pytest
import torch

from oracle import reference_step, taylor
from pinns_fluid_dynamics_b200 import loss_tables, problems
from pinns_fluid_dynamics_b200.engine import assemble_losses, compile_problem

CORONARY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "coronary_geometry.npz")

CASES = {
    "coronary_flow": dict(geometry=CORONARY, PDE=200, BC=800, Vel=50, Pres=0, Test=40, noise_bnd=0.01, noise_fit=0.01),
    "colliding_flow_pressmean": dict(num_pde=200, num_bc=30, num_test=40, num_pres=25),
    "poisson": dict(),
    "poisson_misto": dict(),
    "poiseuille_flow": dict(PDE=200, BC=30, Vel=10, Pres=0, Test=40),
    "colliding_flow": dict(PDE=200, BC=30, Vel=5, Pres=1, Test=40),
    "cavity_steady": dict(PDE=200, BC=30, Vel=20, Pres=1, Test=40, noise_bnd=0.01, noise_fit=0.01),
    "cavity_unsteady": dict(PDE=200, BC=30, IC=20, Vel=3, Pres=1, Test=40, noise_bnd=0.05, noise_fit=0.05, n_times=3),
}


def _both(name, faithful=True, hidden=None, **ref_kw):
    kw = dict(CASES[name])
    if hidden:
        kw["hidden"] = hidden
    data = problems.BUILDERS[name](seed=4, **kw)
    var = reference_step.glorot_uniform_variables(data.dim, data.hidden, data.out_dim, seed=6, bias_std=0.1)
    ref = reference_step.build(data, var, **ref_kw)
    losses, ltest = loss_tables.build_loss_table(data, faithful=faithful)
    cp = compile_problem([tuple(v.shape) for v in var], losses, ltest)
    theta = torch.cat([v.reshape(-1) for v in var]).numpy()
    return ref, cp, theta, losses


@pytest.mark.parametrize("name", sorted(CASES))
def test_taylor_equals_nested(name):
    ref, cp, theta, losses = _both(name)
    vals, total, grad = ref.loss_and_grad()
    out = taylor.loss_and_grad(cp, theta, include_test=True)
    tot2, tr2, te2 = assemble_losses(cp, out[cp.n_params:])
    assert [l.name for l in ref.losses] == [l.name for l in losses]
    assert abs(total - tot2) <= 1e-12 * abs(total)
    assert np.allclose(vals, tr2, rtol=1e-11, atol=1e-300)
    assert np.allclose(ref.test_values(), te2, rtol=1e-11)
    g = grad.numpy()
    assert np.linalg.norm(g - out[:cp.n_params]) <= 1e-11 * np.linalg.norm(g)


@pytest.mark.parametrize("name", ["colliding_flow", "poiseuille_flow"])
def test_in_tape_divergence_variant(name):
    """faithful=False only changes the mass term here when the convection scale is kept: compare the
    mass residual of the corrected table with the oracle's in-tape divergence."""
    ref, cp, theta, losses = _both(name, faithful=False, in_tape_divergence=True)
    vals, _, _ = ref.loss_and_grad()
    out = taylor.loss_and_grad(cp, theta)
    _, tr2, _ = assemble_losses(cp, out[cp.n_params:])
    names = [l.name for l in losses]
    i = names.index("PDE_MASS")
    assert vals[i] > 0 and abs(vals[i] - tr2[i]) <= 1e-11 * vals[i]


def test_coronary_outflow_terms():
    """coronary_flow_steady.py:201-215 calls the model after the tape closed: only the pressure part of the outflow
    condition survives and BCN_v_OUT2 (n = (1, 0), k = 1) is the constant mean(rhs^2) with no gradient -- the flat
    BCN_v_OUT2 log of Test_Case_#123.  faithful=False restores the traction term."""
    ref, cp, theta, losses = _both("coronary_flow")
    vals, _, _ = ref.loss_and_grad()
    names = [l.name for l in losses]
    data = problems.BUILDERS["coronary_flow"](seed=4, **CASES["coronary_flow"])
    i = names.index("BCN_v_OUT2")
    assert abs(vals[i] - np.mean(data.bnd_val[1]["OUT2"] ** 2)) <= 1e-12 * vals[i]
    ref2, cp2, theta2, losses2 = _both("coronary_flow", faithful=False, in_tape_neumann=True)
    vals2, total2, grad2 = ref2.loss_and_grad()
    out = taylor.loss_and_grad(cp2, theta2)
    tot, tr, _ = assemble_losses(cp2, out[cp2.n_params:])
    assert vals2[i] != vals[i]
    assert np.allclose(vals2, tr, rtol=1e-11, atol=1e-300)
    g = grad2.numpy()
    assert np.linalg.norm(g - out[:cp2.n_params]) <= 1e-11 * np.linalg.norm(g)


def test_wide_deep_network():
    """BASELINE config 5 shape (3-128x8-3), few points: the algorithm is width/depth generic."""
    ref, cp, theta, _ = _both("cavity_unsteady", hidden=(128,) * 8)
    _, total, grad = ref.loss_and_grad()
    out = taylor.loss_and_grad(cp, theta)
    tot2, _, _ = assemble_losses(cp, out[cp.n_params:])
    assert cp.n_params == 116483
    assert abs(total - tot2) <= 1e-11 * abs(total)
    g = grad.numpy()
    assert np.linalg.norm(g - out[:cp.n_params]) <= 1e-10 * np.linalg.norm(g)


def test_gradient_matches_finite_differences():
    ref, cp, theta, _ = _both("cavity_steady")
    out = taylor.loss_and_grad(cp, theta)
    rng = np.random.default_rng(0)
    for i in rng.choice(cp.n_params, 6, replace=False):
        e = np.zeros_like(theta); e[i] = 1e-6
        lp, _, _ = assemble_losses(cp, taylor.loss_and_grad(cp, theta + e, with_grad=False)[cp.n_params:])
        lm, _, _ = assemble_losses(cp, taylor.loss_and_grad(cp, theta - e, with_grad=False)[cp.n_params:])
        fd = (lp - lm) / 2e-6
        assert abs(fd - out[i]) <= 1e-5 * max(1.0, abs(out[i]))
