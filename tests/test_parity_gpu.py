"""GPU parity: CUDA loss step (through the facade and the C ABI) vs the float64 oracles on identical
weights and points.  Tolerances are the north-star's: 1e-5 relative on loss values, 1e-4 relative L2
on the parameter gradient."""
import os

import numpy as np
import pytest
import torch

import pinns_fluid_dynamics_b200 as ns
from pinns_fluid_dynamics_b200 import loss_tables, problems
from pinns_fluid_dynamics_b200.engine import assemble_losses

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-5
GRAD_RTOL = 1e-4

CORONARY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "coronary_geometry.npz")

SMALL = {
    "coronary_flow": dict(geometry=CORONARY, PDE=3000, BC=800, Vel=50, Pres=0, Test=1000, noise_bnd=0.01, noise_fit=0.01),
    "colliding_flow_pressmean": dict(),
    "poisson": dict(),
    "poisson_misto": dict(),
    "poiseuille_flow": dict(PDE=1000, BC=100, Vel=10, Pres=0, Test=100),
    "colliding_flow": dict(PDE=1000, BC=100, Vel=5, Pres=1, Test=100),
    "cavity_steady": dict(PDE=1000, BC=100, Vel=100, Pres=1, Test=100, noise_bnd=0.01, noise_fit=0.01),
    "cavity_unsteady": dict(PDE=1000, BC=100, IC=100, Vel=3, Pres=1, Test=100, noise_bnd=0.05, noise_fit=0.05,
                            n_times=4),
}


def _setup(name, kw, seed=1, bias_std=0.1, faithful=True):
    from oracle import reference_step
    data = problems.BUILDERS[name](seed=seed, **kw)
    var = reference_step.glorot_uniform_variables(data.dim, data.hidden, data.out_dim, seed=seed + 10, bias_std=bias_std)
    model = ns.TanhMLP(data.dim, data.hidden, data.out_dim, device="cuda")
    model.set_weights([v.numpy() for v in var])
    losses, ltest = loss_tables.build_loss_table(data, faithful=faithful)
    pb = ns.OptimizationProblem(model.variables, losses, ltest)
    return data, var, model, pb


def _rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def _term_close(v, rv):
    """Per-term check.  A term's value is a mean of squared roots r = (O(1) network output) - (O(1)
    target); FP32 leaves an absolute error of ~1 ulp of the operands in r, i.e. 2*sqrt(value)*1e-7 in
    the value -- that floor matters only for terms that are themselves ~1e-8 (single-point fits)."""
    return abs(v - rv) <= LOSS_RTOL * abs(rv) + 2e-7 * np.sqrt(abs(rv))


@pytest.mark.parametrize("name", list(SMALL))
@pytest.mark.parametrize("bias_std,own_launches", [(0.0, False), (0.1, False), (0.1, True)])
def test_step_matches_reference_restatement(name, bias_std, own_launches, monkeypatch):
    """vs oracle/reference_step.py: nested reverse-mode, one forward per term, the scripts' closures.  ``own_launches``:
    the boundary / fit sets run in kernels of their own derivative order (PINN_NO_PROMOTE=1) instead of riding in the
    collocation launch -- both launch plans must give the reference's numbers."""
    from oracle import reference_step
    if own_launches:
        monkeypatch.setenv("PINN_NO_PROMOTE", "1")
    data, var, model, pb = _setup(name, SMALL[name], bias_std=bias_std)
    total, values, grad = pb.evaluate()
    ref = reference_step.build(data, var)
    ref_values, ref_total, ref_grad = ref.loss_and_grad()
    assert [l.name for l in pb.losses] == [l.name for l in ref.losses]
    assert _rel(total, ref_total) < LOSS_RTOL
    for l, v, rv in zip(pb.losses, values, ref_values):
        if rv == 0.0:
            assert v == 0.0, l.name          # quirk Q1: identically-zero mass residual
        else:
            assert _term_close(v, rv), (l.name, v, rv)
    g = grad.double().cpu().numpy()
    rg = ref_grad.numpy()
    assert np.linalg.norm(g - rg) / np.linalg.norm(rg) < GRAD_RTOL
    # test losses (forward only)
    _, _, test_vals = pb.evaluate_all()
    for v, rv in zip(test_vals, ref.test_values()):
        assert _term_close(v, rv)
    if not own_launches and pb.plan.engine.startswith("fused") and name not in ("colliding_flow_pressmean",):
        pb.plan.loss_and_grad(pb.flat)
        assert pb.plan.last_launch_count() == 2, "small sets ride in the collocation launch: one fused kernel + finalize"


@pytest.mark.parametrize("name", ["colliding_flow", "poiseuille_flow", "cavity_steady", "coronary_flow"])
def test_corrected_residuals_match_taylor_oracle(name):
    """faithful=False (in-tape divergence, -laplacian) vs the Taylor-mode oracle."""
    from oracle import taylor
    data, var, model, pb = _setup(name, SMALL[name], faithful=False)
    total, values, grad = pb.evaluate()
    theta = torch.cat([v.reshape(-1) for v in var]).numpy()
    out = taylor.loss_and_grad(pb.compiled, theta)
    ref_total, ref_vals, _ = assemble_losses(pb.compiled, out[pb.compiled.n_params:])
    assert _rel(total, ref_total) < LOSS_RTOL
    for v, rv in zip(values, ref_vals):
        assert _term_close(v, rv)
    g = grad.double().cpu().numpy()
    rg = out[:pb.compiled.n_params]
    assert np.linalg.norm(g - rg) / np.linalg.norm(rg) < GRAD_RTOL


@pytest.mark.parametrize("n_pde", [1, 15, 16, 17, 33, 4097])
def test_ragged_point_counts(n_pde):
    """chunk tails: point counts that are not a multiple of the 16-point warp chunk."""
    from oracle import taylor
    kw = dict(PDE=n_pde, BC=7, Vel=3, Pres=1, Test=5, noise_bnd=0.01, noise_fit=0.01)
    data, var, model, pb = _setup("cavity_steady", kw)
    total, values, grad = pb.evaluate()
    theta = torch.cat([v.reshape(-1) for v in var]).numpy()
    out = taylor.loss_and_grad(pb.compiled, theta)
    ref_total, ref_vals, _ = assemble_losses(pb.compiled, out[pb.compiled.n_params:])
    assert _rel(total, ref_total) < LOSS_RTOL
    g = grad.double().cpu().numpy()
    rg = out[:pb.compiled.n_params]
    assert np.linalg.norm(g - rg) / np.linalg.norm(rg) < GRAD_RTOL


def test_empty_fit_set_gives_nan_like_reference():
    """Q3: Fit_p over an empty point set is a mean over nothing -> NaN, and it poisons the total."""
    kw = dict(PDE=64, BC=8, Vel=3, Pres=0, Test=5)
    data, var, model, pb = _setup("cavity_steady", kw)
    total, values, grad = pb.evaluate()
    names = [l.name for l in pb.losses]
    assert np.isnan(values[names.index("Fit_p")])
    assert np.isnan(total)
    assert torch.isfinite(grad).all()


def test_large_set_against_taylor_oracle():
    """100k collocation points (Colliding_Flow BASELINE size) vs the Taylor oracle."""
    from oracle import taylor
    kw = dict(PDE=100_000, BC=100, Vel=5, Pres=1, Test=100)
    data, var, model, pb = _setup("colliding_flow", kw, faithful=False)
    total, values, grad = pb.evaluate()
    theta = torch.cat([v.reshape(-1) for v in var]).numpy()
    out = taylor.loss_and_grad(pb.compiled, theta)
    ref_total, _, _ = assemble_losses(pb.compiled, out[pb.compiled.n_params:])
    assert _rel(total, ref_total) < LOSS_RTOL
    g = grad.double().cpu().numpy()
    rg = out[:pb.compiled.n_params]
    assert np.linalg.norm(g - rg) / np.linalg.norm(rg) < GRAD_RTOL


def test_determinism_run_to_run():
    data, var, model, pb = _setup("cavity_steady", SMALL["cavity_steady"])
    a = pb.plan.loss_and_grad(pb.flat).clone()
    b = pb.plan.loss_and_grad(pb.flat).clone()
    assert torch.equal(a, b)


def test_model_forward_matches_oracle():
    from oracle.nisaba_like import KerasMLP
    data, var, model, pb = _setup("cavity_steady", SMALL["cavity_steady"])
    x = torch.rand(1003, 2, dtype=torch.float32)
    y = model(x.cuda()).double().cpu()
    ref = KerasMLP(var)(x.double()).detach()
    assert torch.allclose(y, ref, rtol=0, atol=2e-6)


def test_linearity_in_weights_at_full_size():
    """size-independent property at the BASELINE size (1M collocation points): the gradient of the
    weighted loss is linear in the term weights."""
    kw = dict(PDE=1_000_000, BC=1000, Vel=100, Pres=1, Test=100, noise_bnd=0.01, noise_fit=0.01)
    data = problems.cavity_steady(seed=1, **kw)
    model = ns.TanhMLP(2, [32, 32, 32], 3, device="cuda", seed=3)
    losses, ltest = loss_tables.build_loss_table(data)
    pb1 = ns.OptimizationProblem(model.variables, losses, ltest)
    _, vals1, g1 = pb1.evaluate()
    g1 = g1.clone()
    for l in losses:
        l.weight *= 2.0
    pb2 = ns.OptimizationProblem(model.variables, losses, ltest)
    _, vals2, g2 = pb2.evaluate()
    assert np.allclose(vals1, vals2, rtol=1e-6)
    assert torch.allclose(2.0 * g1, g2, rtol=1e-4, atol=1e-6 * float(g2.abs().max()))


# ---- layered engine (wide / deep networks; BASELINE config 5 shape) ---------------------------------

LAYERED = {
    "unsteady_8x128": ("cavity_unsteady", dict(PDE=777, BC=65, IC=40, Vel=3, Pres=1, Test=50, noise_bnd=0.05,
                                               noise_fit=0.05, n_times=3, hidden=(128,) * 8)),
    "unsteady_2x64": ("cavity_unsteady", dict(PDE=300, BC=33, IC=20, Vel=3, Pres=1, Test=50, n_times=3, hidden=(64,) * 2)),
}


@pytest.mark.parametrize("key", sorted(LAYERED))
def test_layered_engine_matches_taylor_oracle(key):
    from oracle import taylor
    name, kw = LAYERED[key]
    data, var, model, pb = _setup(name, kw)
    assert pb.plan.engine == ("layered_tf32x3" if data.hidden[0] == 128 else "layered_fp32")
    total, values, grad = pb.evaluate()
    theta = torch.cat([v.reshape(-1) for v in var]).numpy()
    out = taylor.loss_and_grad(pb.compiled, theta, include_test=True)
    ref_total, ref_vals, ref_test = assemble_losses(pb.compiled, out[pb.compiled.n_params:])
    assert _rel(total, ref_total) < LOSS_RTOL
    for v, rv in zip(values, ref_vals):
        assert _term_close(v, rv)
    g = grad.double().cpu().numpy()
    rg = out[:pb.compiled.n_params]
    assert np.linalg.norm(g - rg) / np.linalg.norm(rg) < GRAD_RTOL
    _, _, test_vals = pb.evaluate_all()
    for v, rv in zip(test_vals, ref_test):
        assert _term_close(v, rv)


def test_layered_engine_matches_reference_restatement_small():
    """8x128 network against the nested reverse-mode restatement itself (small point counts)."""
    from oracle import reference_step
    kw = dict(PDE=96, BC=20, IC=16, Vel=3, Pres=1, Test=20, noise_bnd=0.05, noise_fit=0.05, n_times=3, hidden=(128,) * 8)
    data, var, model, pb = _setup("cavity_unsteady", kw)
    total, values, grad = pb.evaluate()
    ref = reference_step.build(data, var)
    ref_values, ref_total, ref_grad = ref.loss_and_grad()
    assert _rel(total, ref_total) < LOSS_RTOL
    g = grad.double().cpu().numpy()
    assert np.linalg.norm(g - ref_grad.numpy()) / np.linalg.norm(ref_grad.numpy()) < GRAD_RTOL


def _setup_wide(name, kw, hidden, faithful=True, seed=1):
    """a 2-D script's problem definition on a wide network (tile geometries C = 1, 3, 5 of the tensor-core engine)"""
    from oracle import reference_step
    data = problems.BUILDERS[name](seed=seed, **kw)
    data.hidden = list(hidden)
    var = reference_step.glorot_uniform_variables(data.dim, data.hidden, data.out_dim, seed=seed + 10, bias_std=0.1)
    model = ns.TanhMLP(data.dim, data.hidden, data.out_dim, device="cuda")
    model.set_weights([v.numpy() for v in var])
    losses, ltest = loss_tables.build_loss_table(data, faithful=faithful)
    pb = ns.OptimizationProblem(model.variables, losses, ltest)
    return data, var, model, pb


def _check_against_taylor(pb, var):
    from oracle import taylor
    total, values, grad = pb.evaluate()
    theta = torch.cat([v.reshape(-1) for v in var]).numpy()
    out = taylor.loss_and_grad(pb.compiled, theta, include_test=True)
    ref_total, ref_vals, ref_test = assemble_losses(pb.compiled, out[pb.compiled.n_params:])
    assert _rel(total, ref_total) < LOSS_RTOL, (total, ref_total)
    for v, rv in zip(values, ref_vals):
        assert _term_close(v, rv), (v, rv)
    g = grad.double().cpu().numpy()
    rg = out[:pb.compiled.n_params]
    assert np.linalg.norm(g - rg) / np.linalg.norm(rg) < GRAD_RTOL
    _, _, test_vals = pb.evaluate_all()
    for v, rv in zip(test_vals, ref_test):
        assert _term_close(v, rv)
    return total, grad


@pytest.mark.parametrize("name,kw", [
    ("poiseuille_flow", dict(PDE=500, BC=90, Vel=10, Pres=0, Test=60)),          # Neumann edge: C = 3 tiles
    ("cavity_steady", dict(PDE=1111, BC=70, Vel=30, Pres=1, Test=60, noise_bnd=0.01, noise_fit=0.01)),
])
def test_tensor_core_engine_two_dimensional_problems(name, kw):
    data, var, model, pb = _setup_wide(name, kw, (128,) * 3, faithful=False)
    assert pb.plan.engine == "layered_tf32x3"
    _check_against_taylor(pb, var)


def test_tensor_core_engine_many_batches(monkeypatch):
    """workspace squeezed to one 1680-point batch: 6 batches of collocation points, ragged last tile"""
    monkeypatch.setenv("PINN_TC_WORKSPACE_MB", "1")
    name, kw = LAYERED["unsteady_8x128"]
    kw = dict(kw, PDE=9001)
    data, var, model, pb = _setup(name, kw)
    assert pb.plan.engine == "layered_tf32x3"
    _check_against_taylor(pb, var)


def test_tensor_core_engine_agrees_with_fp32_layered_engine(monkeypatch):
    name, kw = LAYERED["unsteady_8x128"]
    data, var, model, pb = _setup(name, kw)
    t1, v1, g1 = pb.evaluate()
    monkeypatch.setenv("PINN_ENGINE", "layered_fp32")
    data, var, model, pb2 = _setup(name, kw)
    assert pb2.plan.engine == "layered_fp32"
    t2, v2, g2 = pb2.evaluate()
    assert _rel(t1, t2) < LOSS_RTOL
    assert float((g1 - g2).norm() / g2.norm()) < GRAD_RTOL


def test_tensor_core_engine_is_bit_reproducible():
    """no atomics in the layered tensor-core engine: every accumulating kernel adds into its CTAs' own workspace rows and
    the rows are summed in a fixed order, so loss terms and gradient repeat bit for bit"""
    name, kw = LAYERED["unsteady_8x128"]
    data, var, model, pb = _setup(name, kw)
    t1, v1, g1 = pb.evaluate()
    g1 = g1.clone()
    for _ in range(3):
        t2, v2, g2 = pb.evaluate()
        assert t1 == t2 and np.array_equal(np.asarray(v1), np.asarray(v2))
        assert torch.equal(g1, g2)


def test_layered_model_forward():
    from oracle.nisaba_like import KerasMLP
    name, kw = LAYERED["unsteady_8x128"]
    data, var, model, pb = _setup(name, kw)
    x = torch.rand(333, 3, dtype=torch.float32)
    y = model(x.cuda()).double().cpu()
    ref = KerasMLP(var)(x.double()).detach()
    assert torch.allclose(y, ref, rtol=0, atol=5e-6)


# ---- the callers of the path: optimiser rounds --------------------------------

def test_adam_round_follows_float64_oracle():
    """ns.minimize(pb, 'keras', Adam(1e-2), n) (cavity_steady.py:246): device Adam kernel + fused loss step vs the
    float64 restatement driven by the same Keras update rule, same weights and points."""
    from oracle import reference_step
    data, var, model, pb = _setup("cavity_steady", SMALL["cavity_steady"])
    ref = reference_step.build(data, var)
    n_steps = 10
    ns.minimize(pb, "keras", ns.optimizers.Adam(learning_rate=1e-2), num_epochs=n_steps)
    m = [torch.zeros_like(v) for v in ref.variables]
    v2 = [torch.zeros_like(v) for v in ref.variables]
    for t in range(1, n_steps + 1):
        _, _, grad = ref.loss_and_grad()
        off = 0
        with torch.no_grad():
            for i, p in enumerate(ref.variables):
                g = grad[off:off + p.numel()].view_as(p); off += p.numel()
                m[i].mul_(0.9).add_(g, alpha=0.1)
                v2[i].mul_(0.999).addcmul_(g, g, value=0.001)
                step = 1e-2 * (1 - 0.999 ** t) ** 0.5 / (1 - 0.9 ** t)
                p.addcdiv_(m[i], v2[i].sqrt().add_(1e-7), value=-step)
    theta_ref = torch.cat([p.detach().reshape(-1) for p in ref.variables]).numpy()
    theta = pb.flat.double().cpu().numpy()
    assert np.linalg.norm(theta - theta_ref) / np.linalg.norm(theta_ref) < 1e-4
    total, _, _ = pb.evaluate()
    _, ref_total, _ = ref.loss_and_grad()
    assert _rel(total, ref_total) < 1e-3          # 10 steps of 1e-2 amplify the FP32 rounding of the first gradients
    assert pb.history["log"]["iter"][:2] == [0, 10]    # the reference's cadence: every 10 iterations (History_Loss.json)


def test_scipy_bfgs_round_runs_on_the_device_step():
    """ns.minimize(pb, 'scipy', 'BFGS', n) (cavity_steady.py:247): SciPy drives, every evaluation is one CUDA step."""
    data, var, model, pb = _setup("colliding_flow", SMALL["colliding_flow"])
    t0, _, _ = pb.evaluate()
    ns.minimize(pb, "scipy", "BFGS", num_epochs=6)
    t1, _, _ = pb.evaluate()
    assert t1 < t0
    assert pb.history["log_rounds"]["rounds"][-1] == "scipy_BFGS"


def test_tensor_core_engine_linearity_over_several_batches(monkeypatch):
    """size-independent property on the 8x128 network over 3 workspace batches (100 k points, 64 MB workspace ->
    2 520-point batches would be 40; use 1 GB -> 43 680-point batches): the gradient is linear in the term weights."""
    monkeypatch.setenv("PINN_TC_WORKSPACE_MB", "1024")
    kw = dict(PDE=100_000, BC=200, IC=100, Vel=1, Pres=1, Test=50, noise_bnd=0.05, noise_fit=0.05, n_times=4, hidden=(128,) * 8)
    data = problems.cavity_unsteady(seed=1, **kw)
    model = ns.TanhMLP(3, [128] * 8, 3, device="cuda", seed=3)
    losses, ltest = loss_tables.build_loss_table(data)
    pb1 = ns.OptimizationProblem(model.variables, losses, ltest)
    assert pb1.plan.engine == "layered_tf32x3"
    _, vals1, g1 = pb1.evaluate()
    g1 = g1.clone()
    for l in losses:
        l.weight *= 2.0
    pb2 = ns.OptimizationProblem(model.variables, losses, ltest)
    _, vals2, g2 = pb2.evaluate()
    assert np.allclose(vals1, vals2, rtol=2e-5)
    assert float((2.0 * g1 - g2).norm() / g2.norm()) < 1e-4


def test_device_bfgs_round_follows_scipy_round(monkeypatch):
    """the BFGS round with the inverse Hessian on the device vs the same round driven by scipy.optimize.minimize
    (what nisaba calls): same algorithm and line search, so the loss after a few iterations agrees closely even
    though every evaluation is an FP32 device step."""
    def run(mode):
        monkeypatch.setenv("PINN_BFGS", mode)
        data, var, model, pb = _setup("colliding_flow", SMALL["colliding_flow"])
        ns.minimize(pb, "scipy", "BFGS", num_epochs=8)
        return pb.evaluate()[0], pb.flat.double().cpu().numpy()
    t_dev, x_dev = run("device")
    t_ref, x_ref = run("scipy")
    assert _rel(t_dev, t_ref) < 5e-3
    assert np.linalg.norm(x_dev - x_ref) / np.linalg.norm(x_ref) < 1e-3


def test_limited_memory_round_when_the_inverse_hessian_does_not_fit(monkeypatch):
    """P = 116 483 (8x128) would need a 108 GB inverse Hessian: the device driver then runs the L-BFGS two-loop recursion on
    the same line search.  Forced here on the small network: the round must decrease the loss monotonically (strong Wolfe
    steps), stay close to the dense round over the first iterations (identical while fewer than m pairs exist only up to the
    initial scaling), and report its algebra."""
    def run(dense):
        monkeypatch.setenv("PINN_BFGS_DENSE", dense)
        data, var, model, pb = _setup("colliding_flow", SMALL["colliding_flow"])
        t0 = pb.evaluate()[0]
        ns.minimize(pb, "scipy", "BFGS", num_epochs=31)
        return t0, pb.evaluate()[0], pb.last_result, list(pb.history["log"]["loss_global"])
    t0, t_lm, res, hist = run("0")
    assert "L-BFGS" in res.algebra and res.nit == 30
    assert t_lm < 0.5 * t0
    assert all(b <= a * (1 + 1e-6) for a, b in zip(hist, hist[1:]))
    _, t_dense, res_d, _ = run("1")
    assert "dense" in res_d.algebra
    assert t_lm < 3.0 * t_dense and t_dense < 3.0 * t_lm     # same ballpark after 30 iterations


def test_graph_replayed_training_steps_equal_eager_steps(monkeypatch):
    """the CUDA-graph replay of the Adam training step (single GPU) takes exactly the eager steps"""
    def run(flag):
        monkeypatch.setenv("PINN_CUDA_GRAPH", flag)
        data, var, model, pb = _setup("colliding_flow", SMALL["colliding_flow"])
        opt = ns.optimizers.Adam(learning_rate=1e-2)
        for _ in range(12):
            s = pb.training_step(opt)
        return pb.flat.clone(), s.clone(), pb._graph is not None, opt.t
    x_g, s_g, used_g, t_g = run("1")
    x_e, s_e, used_e, t_e = run("0")
    assert used_g and not used_e and t_g == t_e == 12
    assert torch.equal(x_g, x_e)
    assert torch.equal(s_g, s_e)


@pytest.mark.parametrize("shift", [-0.7, 0.7])
def test_abs_mean_term_takes_the_sign_of_the_mean(shift):
    """ns.Loss('PRESS_0', |mean p|) (colliding_flow_pressmean.py:176-179,196): the sign of the mean comes from the forward
    pre-pass inside pinn_loss_and_grad; both signs are forced through the output bias of p."""
    from oracle import reference_step
    data = problems.colliding_flow_pressmean(seed=2, num_pde=500, num_bc=50, num_test=100, num_pres=100)
    var = reference_step.glorot_uniform_variables(data.dim, data.hidden, data.out_dim, seed=5, bias_std=0.1)
    var[-1][2] = float(np.float32(shift))
    model = ns.TanhMLP(data.dim, data.hidden, data.out_dim, device="cuda")
    model.set_weights([v.numpy() for v in var])
    losses, loss_test = loss_tables.build_loss_table(data)
    pb = ns.OptimizationProblem(model.variables, losses, loss_test)
    total, values, grad = pb.evaluate()
    ref = reference_step.build(data, var)
    ref_values, ref_total, ref_grad = ref.loss_and_grad()
    names = [l.name for l in pb.losses]
    i = names.index("PRESS_0")
    assert names == [l.name for l in ref.losses]
    assert ref_values[i] > 0.3 and _rel(values[i], ref_values[i]) < LOSS_RTOL
    assert _rel(total, ref_total) < LOSS_RTOL
    g, rg = grad.double().cpu().numpy(), ref_grad.numpy()
    assert np.linalg.norm(g - rg) / np.linalg.norm(rg) < GRAD_RTOL
    # d/d b_out[2] of w |mean p| is w sign(mean): the last gradient entry carries the sign
    assert np.sign(g[-1]) == np.sign(rg[-1]) == np.sign(shift)
    assert pb.plan.last_launch_count() >= 5      # pre-pass (kernel, finalize, sign) + main pass


def test_tensor_path_fused_engine_agrees_with_fp32_fused_engine(monkeypatch):
    """H = 32: the mma.sync 3xTF32 GEMMs (fused_tf32x3) against the FP32 FFMA2 GEMMs of the same kernel
    (PINN_ENGINE=fused_fp32), two independent implementations of every hidden-layer contraction."""
    data, var, model, pb1 = _setup("cavity_unsteady", SMALL["cavity_unsteady"])
    assert pb1.plan.engine == "fused_tf32x3"
    t1, v1, g1 = pb1.evaluate()
    monkeypatch.setenv("PINN_ENGINE", "fused_fp32")
    losses, loss_test = loss_tables.build_loss_table(data)
    pb2 = ns.OptimizationProblem(model.variables, losses, loss_test)
    assert pb2.plan.engine == "fused_fp32"
    t2, v2, g2 = pb2.evaluate()
    assert _rel(t1, t2) < LOSS_RTOL
    for a, b in zip(v1, v2):
        assert _term_close(a, b)
    assert float((g1 - g2).norm() / g2.norm()) < GRAD_RTOL


def test_prefetched_inputs_restore_the_device_arena():
    """end-to-end path: the double-buffered upload (prefetch on a side stream, device-to-device commit) delivers exactly
    the host arrays -- wipe the device arena, run two prefetch/commit rounds (both pinned host arenas), compare steps."""
    data, var, model, pb = _setup("cavity_steady", SMALL["cavity_steady"])
    t0, v0, g0 = pb.evaluate()
    g0 = g0.clone()
    for _ in range(2):
        pb.plan._arena.zero_()
        pb.plan.prefetch_inputs()
        pb.plan.commit_inputs()
        t1, v1, g1 = pb.evaluate()
        assert t1 == t0 and v1 == v0 and torch.equal(g1, g0)
    pb.plan._arena.zero_()
    pb.plan.upload_inputs()
    t2, _, g2 = pb.evaluate()
    assert t2 == t0 and torch.equal(g2, g0)


# ---- parity at the BASELINE.json sizes (configs 2, 4, 5) and across ranks ---------------------------

def test_poiseuille_at_baseline_size_matches_reference_restatement():
    """BASELINE config 2 at its own size (10 000 collocation points, 4 x 250 boundary points, 10 velocity fit points),
    faithful closures (poiseuille_flow.py:173-254), against the nested reverse-mode restatement."""
    from oracle import reference_step
    kw = dict(PDE=10_000, BC=250, Vel=10, Pres=0, Test=100)
    data, var, model, pb = _setup("poiseuille_flow", kw)
    total, values, grad = pb.evaluate()
    ref = reference_step.build(data, var)
    ref_values, ref_total, ref_grad = ref.loss_and_grad()
    assert _rel(total, ref_total) < LOSS_RTOL
    for l, v, rv in zip(pb.losses, values, ref_values):
        if rv == 0.0:
            assert v == 0.0, l.name
        else:
            assert _term_close(v, rv), (l.name, v, rv)
    g, rg = grad.double().cpu().numpy(), ref_grad.numpy()
    assert np.linalg.norm(g - rg) / np.linalg.norm(rg) < GRAD_RTOL


def test_cavity_steady_20k_matches_reference_restatement():
    """the benchmarked loss table (14 terms, faithful cavity_steady.py:159-231 closures) on 20 000 collocation points
    against the nested reverse-mode restatement: the engine the bench times, checked against the reference's formulation"""
    from oracle import reference_step
    kw = dict(PDE=20_000, BC=1000, Vel=100, Pres=1, Test=100, noise_bnd=0.01, noise_fit=0.01)
    data, var, model, pb = _setup("cavity_steady", kw)
    assert pb.plan.engine == "fused_tcgen05"
    total, values, grad = pb.evaluate()
    ref = reference_step.build(data, var)
    ref_values, ref_total, ref_grad = ref.loss_and_grad()
    assert _rel(total, ref_total) < LOSS_RTOL
    for l, v, rv in zip(pb.losses, values, ref_values):
        assert _term_close(v, rv), (l.name, v, rv)
    g, rg = grad.double().cpu().numpy(), ref_grad.numpy()
    assert np.linalg.norm(g - rg) / np.linalg.norm(rg) < GRAD_RTOL


def test_cavity_steady_at_baseline_size_matches_taylor_oracle():
    """BASELINE config 4 at its own size -- the configuration bench.py times: 1 000 000 collocation points, 4 x 1000
    boundary points, 100 + 1 fit points, 14 terms -- against the float64 Taylor-mode oracle (chunked over the points)."""
    from oracle import taylor
    kw = dict(PDE=1_000_000, BC=1000, Vel=100, Pres=1, Test=100, noise_bnd=0.01, noise_fit=0.01)
    data, var, model, pb = _setup("cavity_steady", kw)
    assert pb.plan.engine == "fused_tcgen05"
    total, values, grad = pb.evaluate()
    theta = torch.cat([v.reshape(-1) for v in var]).numpy()
    out = taylor.loss_and_grad(pb.compiled, theta, chunk=131072)
    ref_total, ref_vals, _ = assemble_losses(pb.compiled, out[pb.compiled.n_params:])
    assert _rel(total, ref_total) < LOSS_RTOL
    for l, v, rv in zip(pb.losses, values, ref_vals):
        assert _term_close(v, rv), (l.name, v, rv)
    g, rg = grad.double().cpu().numpy(), out[:pb.compiled.n_params]
    assert np.linalg.norm(g - rg) / np.linalg.norm(rg) < GRAD_RTOL


_WIDE_200K = {}


def _wide_200k_reference():
    """float64 Taylor-mode oracle of the 8x128 unsteady network on 200 000 collocation points (computed once)"""
    from oracle import reference_step, taylor
    if not _WIDE_200K:
        kw = dict(PDE=200_000, BC=500, IC=500, Vel=1, Pres=1, Test=50, noise_bnd=0.05, noise_fit=0.05, n_times=4, hidden=(128,) * 8)
        data = problems.cavity_unsteady(seed=1, **kw)
        var = reference_step.glorot_uniform_variables(data.dim, data.hidden, data.out_dim, seed=11, bias_std=0.1)
        _WIDE_200K.update(data=data, var=var)
    return _WIDE_200K


@pytest.mark.parametrize("workspace_mb", [None, 2048])
def test_tensor_core_engine_200k_points_matches_taylor_oracle(monkeypatch, workspace_mb):
    """BASELINE config 5 network (3-128x8-3) on 200 000 space-time collocation points against the Taylor oracle, with the
    default 16 GiB workspace (one batch) and with a 2 GiB workspace (3 batches of 83 160 points, ragged last batch)."""
    from oracle import taylor
    if workspace_mb is not None:
        monkeypatch.setenv("PINN_TC_WORKSPACE_MB", str(workspace_mb))
    else:
        monkeypatch.delenv("PINN_TC_WORKSPACE_MB", raising=False)
    ref = _wide_200k_reference()
    data, var = ref["data"], ref["var"]
    model = ns.TanhMLP(data.dim, data.hidden, data.out_dim, device="cuda")
    model.set_weights([v.numpy() for v in var])
    losses, ltest = loss_tables.build_loss_table(data)
    pb = ns.OptimizationProblem(model.variables, losses, ltest)
    assert pb.plan.engine.startswith("layered_tf32x3") or pb.plan.engine.startswith("fused_wide")
    total, values, grad = pb.evaluate()
    if "out" not in ref:
        theta = torch.cat([v.reshape(-1) for v in var]).numpy()
        ref["out"] = taylor.loss_and_grad(pb.compiled, theta, chunk=16384)
    out = ref["out"]
    ref_total, ref_vals, _ = assemble_losses(pb.compiled, out[pb.compiled.n_params:])
    assert _rel(total, ref_total) < LOSS_RTOL
    for l, v, rv in zip(pb.losses, values, ref_vals):
        assert _term_close(v, rv), (l.name, v, rv)
    g, rg = grad.double().cpu().numpy(), out[:pb.compiled.n_params]
    assert np.linalg.norm(g - rg) / np.linalg.norm(rg) < GRAD_RTOL


def test_two_nccl_ranks_match_the_unsharded_oracle():
    """N > 1 on the device: two ranks (torchrun) evaluate their shards of every point set, one all-reduce joins them (peer-memory
    kernel for the 3x32 network, NCCL for 8x128), and the global loss terms and gradient must match the float64 oracle of the
    WHOLE problem (both engines); then the two all-reduce paths are run side by side over 310 Adam steps.
    Skipped on a single-GPU box; tests/test_distributed_gloo.py covers the same host logic on CPU."""
    import re
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(root, "tools", "multi_gpu_check.py")]
    env = dict(os.environ, PINN_P2P_COMPARE="1", PINN_P2P_COMPARE_PDE="20000")
    env.pop("PINN_P2P_ALLREDUCE", None)
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if "world=2" in l and "total rel err" in l]
    assert len(lines) == 2, res.stdout
    for l in lines:
        m = re.search(r"total rel err ([0-9.e+-]+), worst term ([0-9.e+-]+), grad rel L2 ([0-9.e+-]+)", l)
        assert m, l
        assert float(m.group(1)) < LOSS_RTOL and float(m.group(2)) < 2e-5 and float(m.group(3)) < GRAD_RTOL, l
    # the two all-reduce paths (ncclAllReduce in the step's graph / the one-shot peer-memory kernel): both really used, no wait timed
    # out, every rank holds bit-identical parameters after 310 Adam steps, and the two runs agree
    paths = [l for l in res.stdout.splitlines() if l.startswith("allreduce path requested")]
    assert len(paths) == 2, res.stdout
    assert "requested nccl, used nccl" in paths[0] and "requested p2p, used p2p" in paths[1], paths
    assert all("timeouts 0" in l and "bit-identical: True" in l for l in paths), paths
    m = re.search(r"relative difference of the parameters ([0-9.e+-]+)", res.stdout)
    assert m and float(m.group(1)) < 1e-5, res.stdout


@pytest.mark.parametrize("n_pde", [700, 4097])
def test_more_than_four_terms_on_one_point_set(n_pde):
    """A point set with eight train terms (the ABI's maximum per set) and a test set with six: the fused kernels keep the first four sums of a set in registers
    and take another path for the rest; the four threads of a point share the terms in groups of four (two groups here).  Compared term by term with the Taylor oracle."""
    from oracle import reference_step, taylor
    from pinns_fluid_dynamics_b200 import residuals as R
    from pinns_fluid_dynamics_b200.api import LossMeanSquares as LMS
    from pinns_fluid_dynamics_b200.residuals import PointSet
    kw = dict(PDE=n_pde, BC=50, Vel=20, Pres=1, Test=333, noise_bnd=0.01, noise_fit=0.01)
    data = problems.cavity_steady(seed=5, **kw)
    rng = np.random.default_rng(7)
    pde, test = PointSet(data.x_pde, "PDE"), PointSet(data.x_test, "Test")
    nv, npr = data.norm_vel, data.norm_pre
    tgt = [rng.normal(size=data.x_pde.shape[0]) for _ in range(3)]
    mom = lambda k, vxx: R.momentum(pde, k, nv, npr, conv_scale=nv, visc_xx=vxx, visc_yy=-1.0)
    losses = [LMS("mass", lambda: R.mass(pde), weight=10.0),
              LMS("momu", lambda: mom(0, 1.0), weight=1.0),
              LMS("momv", lambda: mom(1, 1.0), weight=1.0),
              LMS("fit_u", lambda: R.dirichlet(pde, 0, tgt[0]), weight=0.5),
              LMS("fit_v", lambda: R.dirichlet(pde, 1, tgt[1]), weight=2.0),
              LMS("fit_p", lambda: R.dirichlet(pde, 2, tgt[2]), weight=0.25),
              LMS("momu_lap", lambda: mom(0, -1.0), weight=0.1),
              LMS("momv_lap", lambda: mom(1, -1.0), weight=0.3)]
    ltest = [LMS(f"t{i}", lambda i=i: R.dirichlet(test, i % 3, data.sol_test[i % 3] * (1.0 + i))) for i in range(6)]
    var = reference_step.glorot_uniform_variables(data.dim, data.hidden, data.out_dim, seed=21, bias_std=0.1)
    model = ns.TanhMLP(data.dim, data.hidden, data.out_dim, device="cuda")
    model.set_weights([v.numpy() for v in var])
    pb = ns.OptimizationProblem(model.variables, losses, ltest)
    total, values, grad = pb.evaluate()
    theta = torch.cat([v.reshape(-1) for v in var]).numpy()
    out = taylor.loss_and_grad(pb.compiled, theta)
    ref_total, ref_vals, _ = assemble_losses(pb.compiled, out[pb.compiled.n_params:])
    assert _rel(total, ref_total) < LOSS_RTOL
    assert len(values) == 8
    for v, rv in zip(values, ref_vals):
        assert _term_close(v, rv)
    g = grad.double().cpu().numpy()
    rg = out[:pb.compiled.n_params]
    assert np.linalg.norm(g - rg) / np.linalg.norm(rg) < GRAD_RTOL
    _, all_vals, test_vals = pb.evaluate_all()                 # forward-only kernel: train and test terms
    out_all = taylor.loss_and_grad(pb.compiled, theta, with_grad=False, include_test=True)
    _, ref_all, ref_test = assemble_losses(pb.compiled, out_all[pb.compiled.n_params:])
    assert len(test_vals) == 6
    for v, rv in zip(list(all_vals) + list(test_vals), list(ref_all) + list(ref_test)):
        assert _term_close(v, rv)
