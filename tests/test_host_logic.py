"""Host-side logic: loss-table compilation, sharding, loss assembly, the nisaba-shaped facade and its
history bookkeeping.  The evaluator is the float64 Taylor oracle injected through ``engine_factory``
(test-only hook); the product evaluator is CUDA-only."""
import json
import os

import numpy as np
import pytest
import torch

import pinns_fluid_dynamics_b200 as ns
from oracle.taylor import TaylorEngine
from pinns_fluid_dynamics_b200 import _capi, loss_tables, problems
from pinns_fluid_dynamics_b200.engine import (assemble_losses, compile_problem, mlp_shape_from_variables, param_count,
                                              shard_bounds)

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _problem(name="cavity_steady", **kw):
    base = dict(PDE=64, BC=12, Vel=6, Pres=1, Test=9, noise_bnd=0.01, noise_fit=0.01)
    base.update(kw)
    data = problems.BUILDERS[name](seed=2, **base)
    model = ns.TanhMLP(data.dim, data.hidden, data.out_dim, device="cpu", seed=5)
    losses, ltest = loss_tables.build_loss_table(data)
    return data, model, losses, ltest


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 5, 16, 1000, 1_000_003):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [e - s for s, e in b]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def test_param_counts_match_survey():
    assert param_count(2, 20, 3, 1) == 921 and param_count(2, 32, 3, 3) == 2307
    assert param_count(3, 32, 3, 3) == 2339 and param_count(3, 128, 8, 3) == 116483
    assert mlp_shape_from_variables([(2, 32), (32,), (32, 32), (32,), (32, 32), (32,), (32, 3), (3,)]) == (2, 32, 3, 3)
    with pytest.raises(ValueError):
        mlp_shape_from_variables([(2, 32), (32,), (31, 32), (32,), (32, 3), (3,)])


def test_terms_on_one_point_set_are_fused():
    data, model, losses, ltest = _problem()
    cp = compile_problem(model.shapes, losses, ltest)
    by_name = {cs.pointset.name: cs for cs in cp.sets}
    assert [t.name for t in by_name["PDE"].terms] == ["PDE_MASS", "PDE_MOMU", "PDE_MOMV"]
    assert by_name["PDE"].deriv_order == 2
    assert [t.name for t in by_name["TOP"].terms] == ["BCD_u_y1", "BCD_v_y1"] and by_name["TOP"].deriv_order == 0
    assert [t.name for t in by_name["Test"].terms] == ["u_test", "v_test", "p_test"]
    # output order is the table's, train terms first
    assert [t.name for t in sorted(cp.terms, key=lambda t: t.out_index)] == [l.name for l in losses + ltest]
    assert len(cp.sets) == 8    # PDE, 4 edges, Vel, Pres, Test: 17 reference forwards -> 8 fused passes


def test_neumann_needs_first_derivatives_and_q1_term_is_dropped():
    data, model, losses, ltest = _problem("poiseuille_flow", Pres=0)
    cp = compile_problem(model.shapes, losses, ltest)
    by_name = {cs.pointset.name: cs for cs in cp.sets}
    assert by_name["DX"].deriv_order == 1
    mass = [t for t in cp.terms if t.name == "PDE_MASS"][0]
    assert mass.out_index == -1 and all(t.name != "PDE_MASS" for cs in cp.sets for t in cs.terms)
    total, vals, _ = assemble_losses(cp, np.ones(cp.n_out_terms))
    assert vals[0] == 0.0


def test_facade_rejects_closures_and_foreign_variables():
    data, model, losses, ltest = _problem()
    with pytest.raises(TypeError):
        ns.LossMeanSquares("x", lambda: torch.zeros(3))
    with pytest.raises(NotImplementedError):
        ns.Loss("PRESS_0", lambda: 0.0)
    from pinns_fluid_dynamics_b200 import residuals as R
    ps = R.PointSet(np.zeros((4, 2)))
    with pytest.raises(NotImplementedError):       # ns.Loss takes the |mean| reduction only
        ns.Loss("PRESS_0", lambda: R.dirichlet(ps, 2))
    l = ns.Loss("PRESS_0", lambda: R.mean_value(ps, 2), normalization=1e0, weight=1e-2, non_negative=True)
    assert (l.weight, l.normalization, l.non_negative, l.form.reduction) == (1e-2, 1.0, True, "abs_mean")
    loose = [v.clone() for v in model.variables]
    with pytest.raises(ValueError):
        ns.OptimizationProblem(loose, losses, ltest, engine_factory=TaylorEngine)


def test_product_evaluator_refuses_to_run_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    data, model, losses, ltest = _problem()
    with pytest.raises(_capi.PinnLibraryError):
        ns.OptimizationProblem(model.variables, losses, ltest)
    with pytest.raises(_capi.PinnLibraryError):
        model(np.zeros((4, 2), dtype=np.float32))


def test_history_schema_and_bookkeeping_match_reference(tmp_path):
    data, model, losses, ltest = _problem()
    pb = ns.OptimizationProblem(model.variables, losses, ltest, callbacks=[], engine_factory=TaylorEngine)
    hist_file = tmp_path / "History_Loss.json"
    pb.callbacks.append(ns.utils.HistoryPlotCallback(frequency=100, gui=False, filename=str(tmp_path / "x.png"),
                                                     filename_history=str(hist_file)))
    ns.minimize(pb, "keras", ns.optimizers.Adam(learning_rate=1e-2), num_epochs=100)
    ns.minimize(pb, "scipy", "BFGS", num_epochs=25)
    h = pb.history
    ref = json.load(open(os.path.join(GOLD, "history_cavity_steady.json")))
    assert set(h) == {"log", "losses", "losses_test", "log_rounds"} and set(h["log"]) == set(ref["log"])
    assert list(h["losses"]) == list(ref["losses"]) and list(h["losses_test"]) == list(ref["losses_test"])
    for n in h["losses"]:
        assert {k: h["losses"][n][k] for k in ("weight", "non_negative", "display_sqrt")} == \
               {k: ref["losses"][n][k] for k in ("weight", "non_negative", "display_sqrt")}
    assert h["log"]["iter"][:13] == ref["log"]["iter"][:13]
    assert h["log"]["iter_round"][:13] == ref["log"]["iter_round"][:13]
    assert h["log"]["round"][:13] == ref["log"]["round"][:13]
    assert h["log_rounds"] == ref["log_rounds"]
    assert h["log"]["loss_global"][10] == h["log"]["loss_global"][11]
    tot = sum(d["weight"] * np.asarray(d["log"]) for d in h["losses"].values())
    assert np.allclose(tot, h["log"]["loss_global"], rtol=1e-12)
    assert h["log"]["loss_global"][10] < h["log"]["loss_global"][0]      # Adam made progress
    assert h["log"]["loss_global"][-1] <= h["log"]["loss_global"][11]    # and so did BFGS
    saved = ns.utils.load_json(str(hist_file))
    assert set(saved) == set(h)
    pb.save_history(str(tmp_path / "h2.json"))
    assert json.load(open(tmp_path / "h2.json"))["log_rounds"]["rounds"] == ["keras_Adam", "scipy_BFGS"]


def test_model_json_and_weights_export(tmp_path):
    data, model, losses, ltest = _problem()
    j = json.loads(model.to_json())
    assert [l["config"].get("units") for l in j["config"]["layers"][1:]] == [32, 32, 32, 3]
    model.save_weights(str(tmp_path / "Weights.npz"))
    w = np.load(tmp_path / "Weights.npz")
    assert w["dense_0/kernel:0"].shape == (2, 32) and w["dense_3/bias:0"].shape == (3,)
    m2 = ns.TanhMLP(2, [32, 32, 32], 3, device="cpu")
    m2.set_weights(model.get_weights())
    assert torch.equal(m2.flat, model.flat)
    lim = np.sqrt(6.0 / (2 + 32))
    assert float(model.variables[0].abs().max()) <= lim and float(model.variables[1].abs().max()) == 0.0


def test_host_objective_matches_evaluate():
    """``evaluate_host`` (the objective SciPy-style drivers call) == ``evaluate`` at the same parameters; without CUDA it
    takes the plain path through the injected engine."""
    data, model, losses, ltest = _problem()
    pb = ns.OptimizationProblem(model.variables, losses, ltest, engine_factory=TaylorEngine)
    theta = pb.flat.detach().double().numpy().copy()
    theta += 1e-3 * np.random.default_rng(0).standard_normal(theta.shape)
    f, g = pb.evaluate_host(theta)
    assert np.allclose(pb.flat.double().numpy(), theta.astype(np.float32).astype(np.float64))   # parameters were installed
    total, _, grad = pb.evaluate()
    assert f == total and np.array_equal(g, grad.double().numpy())
