#!/usr/bin/env python
"""Colliding_Flow on the B200 framework: the top-level flow of Examples/Colliding_Flow/colliding_flow.py
(options file :37-58, point sets :88-150, loss table :209-228, OptimizationProblem + HistoryPlotCallback :239-241,
Adam x100 then BFGS x epochs :242-244, Model.json / weights / history / recap :246-379) with the nisaba names served by
``pinns_fluid_dynamics_b200``.  Plots are out of scope.

    python examples/colliding_flow.py [--epochs N] [--out DIR]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pinns_fluid_dynamics_b200 as ns
from pinns_fluid_dynamics_b200 import loss_tables, options, problems

ap = argparse.ArgumentParser()
ap.add_argument("--epochs", type=int, default=None, help="override TRAINING EPOCHS of simulation_options.txt")
ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "Test_Case_colliding_flow"))
args = ap.parse_args()

# %% Options -- the reference's 20-line positional file; written with the checked-in values when absent
opt_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "simulation_options.txt")
if not os.path.exists(opt_path):
    o = options.SimulationOptions(epochs=10000, noise_factor_fit=0.0, noise_factor_bnd=0.0)
    o.n_pts.update({"PDE": 1000, "BC": 100, "IC": 100, "Vel": 5, "Pres": 1, "Test": 10000})
    options.write_simulation_options(opt_path, o)
opt = options.read_simulation_options(opt_path)
epochs = opt.epochs if args.epochs is None else args.epochs

# %% Problem definition, loss table, model (Keras layout, GlorotUniform / zeros)
data = problems.colliding_flow(options=opt, seed=1)
model = ns.TanhMLP(data.dim, data.hidden, data.out_dim, device="cuda", seed=1)
losses, loss_test = loss_tables.build_loss_table(data)          # faithful: PDE_MASS logs 0 (quirk Q1)

# %% Training
os.makedirs(args.out, exist_ok=True)
pb = ns.OptimizationProblem(model.variables, losses, loss_test, callbacks=[])
pb.callbacks.append(ns.utils.HistoryPlotCallback(frequency=100, gui=False,
                                                 filename=os.path.join(args.out, "History_Loss.png"),
                                                 filename_history=os.path.join(args.out, "History_Loss.json")))
t0 = time.perf_counter()
ns.minimize(pb, "keras", ns.optimizers.Adam(learning_rate=1e-2), num_epochs=100)
torch.cuda.synchronize()
t1 = time.perf_counter()
ns.minimize(pb, "scipy", "BFGS", num_epochs=epochs)
torch.cuda.synchronize()
t2 = time.perf_counter()

# %% Saving
with open(os.path.join(args.out, "Model.json"), "w") as fh:
    fh.write(model.to_json())
model.save_weights(os.path.join(args.out, "Weights.npz"))
pb.save_history(os.path.join(args.out, "History_Loss.json"))
total, train_vals, test_vals = pb.evaluate_all()
recap = {"engine": pb.plan.engine, "adam_seconds": t1 - t0, "bfgs_seconds": t2 - t1, "bfgs_iterations": pb.history["log"]["iter"][-1] - 100,
         "loss_global": total, "losses": {l.name: v for l, v in zip(pb.losses, train_vals)},
         "losses_test": {l.name: v for l, v in zip(pb.losses_test, test_vals)}}
options.write_recap(os.path.join(args.out, "Test_Options.txt"), "Colliding_Flow", opt)     # colliding_flow.py:362-379
with open(os.path.join(args.out, "Run_Summary.json"), "w") as fh:
    json.dump({"epochs": epochs, "n_pts": opt.n_pts, **recap}, fh, indent=2)
print(json.dumps(recap, indent=1))
