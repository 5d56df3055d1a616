"""Shared top-level flow of the example scripts: options file -> problem -> loss table -> OptimizationProblem with the history
callback -> Adam round -> quasi-Newton round -> Model.json / Weights.h5 / History_Loss.json / Test_Options.txt, the sequence
every reference script runs (e.g. cavity_steady.py:37-58, 236-252, 395-412).  Plots are out of scope; the fields a script would
plot are saved as arrays instead."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pinns_fluid_dynamics_b200 as ns  # noqa: E402
from pinns_fluid_dynamics_b200 import loss_tables, options  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def read_or_write_options(case: str, defaults: dict, counts: dict):
    """The reference's 20-line positional ``simulation_options.txt`` (one per problem folder): written with the checked-in
    values of that problem when absent, then read back through the same parser a user's file goes through."""
    path = os.path.join(HERE, f"simulation_options_{case}.txt")
    if not os.path.exists(path):
        o = options.SimulationOptions(**defaults)
        o.n_pts.update(counts)
        options.write_simulation_options(path, o)
    return options.read_simulation_options(path)


def train_and_save(case: str, data, model, opt, epochs: int, out_dir: str, adam_epochs: int = 100, method: str = "BFGS", extra=None):
    """OptimizationProblem + HistoryPlotCallback, ``adam_epochs`` of Adam(1e-2), ``epochs`` of the SciPy method, then the files
    the reference writes into its Test_Case folder."""
    losses, loss_test = loss_tables.build_loss_table(data)
    os.makedirs(out_dir, exist_ok=True)
    pb = ns.OptimizationProblem(model.variables, losses, loss_test, callbacks=[])
    pb.callbacks.append(ns.utils.HistoryPlotCallback(frequency=100, gui=False, filename=os.path.join(out_dir, "Loss_Trend_Full.png"),
                                                     filename_history=os.path.join(out_dir, "History_Loss.json")))
    t0 = time.perf_counter()
    ns.minimize(pb, "keras", ns.optimizers.Adam(learning_rate=1e-2), num_epochs=adam_epochs)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    ns.minimize(pb, "scipy", method, num_epochs=epochs)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    with open(os.path.join(out_dir, "Model.json"), "w") as fh:
        fh.write(model.to_json())
    model.save_weights(os.path.join(out_dir, "Weights.h5"))
    pb.save_history(os.path.join(out_dir, "History_Loss.json"))
    if opt is not None:
        options.write_recap(os.path.join(out_dir, "Test_Options.txt"), case, opt)
    total, train_vals, test_vals = pb.evaluate_all()
    res = getattr(pb, "last_result", None)
    recap = {"case": case, "engine": pb.plan.engine, "adam_seconds": t1 - t0, "quasi_newton_seconds": t2 - t1,
             "quasi_newton_iterations": int(res.nit) if res is not None and "nit" in res else None,
             "loss_evaluations": int(res.nfev) if res is not None and "nfev" in res else None,
             "loss_global": total, "losses": {l.name: v for l, v in zip(pb.losses, train_vals)},
             "losses_test": {l.name: v for l, v in zip(pb.losses_test, test_vals)}}
    if extra:
        recap.update(extra)
    with open(os.path.join(out_dir, "Run_Summary.json"), "w") as fh:
        json.dump(recap, fh, indent=2)
    print(json.dumps(recap, indent=1))
    return pb, recap
