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
    # the dataset paths of Keras' Weights.h5 (Test_Case folders: dense_N/dense_N/kernel:0)
    assert w["dense/dense/kernel:0"].shape == (2, 32) and w["dense_3/dense_3/bias:0"].shape == (3,)
    m2 = ns.TanhMLP(2, [32, 32, 32], 3, device="cpu")
    m2.set_weights(model.get_weights())
    assert torch.equal(m2.flat, model.flat)
    lim = np.sqrt(6.0 / (2 + 32))
    assert float(model.variables[0].abs().max()) <= lim and float(model.variables[1].abs().max()) == 0.0


@pytest.mark.parametrize("name", ["Weights.h5", "Weights.npz"])
def test_weights_round_trip(tmp_path, name):
    """model.save_weights / load_weights (cavity_steady.py:252): the HDF5 file carries Keras' dataset paths, is read
    back by the same reader that reads the reference's Weights.h5, and restores the flat vector bit for bit."""
    from pinns_fluid_dynamics_b200 import h5lite
    model = ns.TanhMLP(3, [32] * 3, 3, device="cpu", seed=7)
    with torch.no_grad():
        model.variables[1].uniform_(-0.1, 0.1)          # non-zero biases so their order matters
        model.variables[5].uniform_(-0.1, 0.1)
    path = str(tmp_path / name)
    model.save_weights(path)
    if name.endswith(".h5"):
        f = h5lite.H5File(path)
        assert f["dense_2/dense_2/kernel:0"].shape == (32, 32) and f["dense/dense/bias:0"].shape == (32,)
        assert f["dense_3/dense_3/kernel:0"].dtype == np.float64      # the reference's files are float64 (Model.json)
    other = ns.TanhMLP(3, [32] * 3, 3, device="cpu", seed=8)
    assert not torch.equal(other.flat, model.flat)
    other.load_weights(path)
    assert torch.equal(other.flat, model.flat)
    with pytest.raises(ValueError):
        ns.TanhMLP(2, [32] * 3, 3, device="cpu").load_weights(path)     # wrong architecture


def test_h5_writer_many_layers_and_nested_groups(tmp_path):
    """the writer's group B-tree over more links than one symbol-table node holds (8x128 network: 9 layers + nesting)"""
    from pinns_fluid_dynamics_b200 import h5lite
    rng = np.random.default_rng(0)
    data = {f"g{i}/sub/x{j}": rng.standard_normal((i + 1, j + 2)).astype(np.float32 if j % 2 else np.float64)
            for i in range(11) for j in range(3)}
    data["top"] = np.arange(5, dtype=np.int32)
    h5lite.write_h5(str(tmp_path / "t.h5"), data)
    f = h5lite.H5File(str(tmp_path / "t.h5"))
    for k, v in data.items():
        got = f[k]
        assert got.dtype == v.dtype and np.array_equal(got, v), k
    m = ns.TanhMLP(3, [128] * 8, 3, device="cpu", seed=1)
    m.save_weights(str(tmp_path / "w.h5"))
    m2 = ns.TanhMLP(3, [128] * 8, 3, device="cpu", seed=2)
    m2.load_weights(str(tmp_path / "w.h5"))
    assert torch.equal(m.flat, m2.flat)


def test_unsteady_fem_files_are_read_like_the_script(tmp_path):
    """cavity_unsteady.py:103-113: one HDF5 file per time step (VisualisationVector/0 velocity, /1 pressure), the pressure
    of every step minus ITS OWN mean, steps concatenated in time order over the (t, y, x) grid."""
    from pinns_fluid_dynamics_b200 import h5lite
    n_times, n = 3, 101 * 101
    rng = np.random.default_rng(5)
    vel = rng.standard_normal((n_times, n, 3))
    pre = rng.standard_normal((n_times, n)) + np.array([10.0, -3.0, 0.5])[:, None]
    for t in range(n_times):
        h5lite.write_h5(str(tmp_path / f"navier-stokes_SI_cavity_unsteady_{t:05d}.h5"),
                        {"VisualisationVector/0": vel[t], "VisualisationVector/1": pre[t][:, None],
                         "Mesh/0/mesh/geometry": np.zeros((n, 2))})
    u, v, p = problems.load_unsteady_fem_fields(str(tmp_path), n_times)
    assert np.array_equal(u, vel[:, :, 0].reshape(-1)) and np.array_equal(v, vel[:, :, 1].reshape(-1))
    assert np.allclose(p.reshape(n_times, n).mean(axis=1), 0.0, atol=1e-12)
    assert np.allclose(p, (pre - pre.mean(axis=1, keepdims=True)).reshape(-1))
    kw = dict(PDE=300, BC=20, IC=20, Vel=10, Pres=5, Test=10, noise_bnd=0.0, noise_fit=0.0, n_times=n_times)
    d = problems.cavity_unsteady(seed=2, fem_fields=str(tmp_path), **kw)
    assert d.norm_vel == pytest.approx(max(np.ptp(u), np.ptp(v))) and d.norm_pre == pytest.approx(np.ptp(p))
    # the fitting targets are the file values at the sampled grid rows over the normalisation (cavity_unsteady.py:115-123,163)
    d2 = problems.cavity_unsteady(seed=2, fem_fields=(u, v, p), **kw)
    assert np.array_equal(d.x_pde, d2.x_pde)
    for a, b in zip(jax_free_leaves(d), jax_free_leaves(d2)):
        assert np.array_equal(a, b)
    with pytest.raises(FileNotFoundError):
        problems.load_unsteady_fem_fields(str(tmp_path), n_times + 1)
    with pytest.raises(ValueError):
        problems.cavity_unsteady(seed=2, fem_fields=(u[:-1], v[:-1], p[:-1]), **kw)


def jax_free_leaves(d):
    """every numpy array a ProblemData holds (point sets and targets), in a fixed order"""
    out = []
    for key in sorted(vars(d)):
        val = getattr(d, key)
        if isinstance(val, np.ndarray):
            out.append(val)
        elif isinstance(val, dict):
            out += [val[k] for k in sorted(val) if isinstance(val[k], np.ndarray)]
        elif isinstance(val, (list, tuple)):
            for item in val:
                if isinstance(item, np.ndarray):
                    out.append(item)
                elif isinstance(item, dict):
                    out += [item[k] for k in sorted(item) if isinstance(item[k], np.ndarray)]
    return out


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
