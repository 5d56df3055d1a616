#!/usr/bin/env python
"""Coronary_Flow on the B200 framework: the top-level flow of Examples/Coronary_Flow/coronary_flow_steady.py
(options file :37-58, mesh nodes / labelled boundary points / reference fields :92-136, loss table :217-246,
Adam x100 then BFGS x epochs :254-255, Model.json / weights :257-260, solution at the mesh nodes :297-301, recap :366-383).

    python examples/coronary_flow.py [--epochs N] [--out DIR] [--geometry coronary_geometry.npz | --fem FEM.h5 BPOINTS.npy]

Without DataGeneration output the geometry fixture of tests/golden (mesh nodes of coroParam.msh, bpoints.npy and the
u/v/p arrays of the reference's sol_pinn.h5 as the field to fit) is used.  Plots are out of scope.
"""
import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import numpy as np
import torch

import pinns_fluid_dynamics_b200 as ns
from pinns_fluid_dynamics_b200 import loss_tables, options, problems

ap = argparse.ArgumentParser()
ap.add_argument("--epochs", type=int, default=None, help="override TRAINING EPOCHS of the options file")
ap.add_argument("--out", default=os.path.join(HERE, "Test_Case_coronary_flow"))
ap.add_argument("--geometry", default=os.path.join(os.path.dirname(HERE), "tests", "golden", "coronary_geometry.npz"))
ap.add_argument("--fem", nargs=2, metavar=("FEM_H5", "BPOINTS_NPY"), default=None)
args = ap.parse_args()

# %% Options: Examples/Coronary_Flow/simulation_options.txt values
opt = options.SimulationOptions(epochs=30000, noise_factor_fit=0.01, noise_factor_bnd=0.01)
opt.n_pts.update({"PDE": 3000, "BC": 800, "IC": 0, "Vel": 50, "Pres": 0, "Test": 1000})
epochs = opt.epochs if args.epochs is None else args.epochs

# %% Problem definition, loss table, model
geometry = problems.load_coronary_geometry(tuple(args.fem) if args.fem else args.geometry)
data = problems.coronary_flow(geometry, options=opt, seed=1)
model = ns.TanhMLP(data.dim, data.hidden, data.out_dim, device="cuda", seed=1)
losses, loss_test = loss_tables.build_loss_table(data)     # faithful: outflow terms keep only their pressure part

# %% Training
os.makedirs(args.out, exist_ok=True)
pb = ns.OptimizationProblem(model.variables, losses, loss_test, callbacks=[])
pb.callbacks.append(ns.utils.HistoryPlotCallback(frequency=100, gui=False,
                                                 filename=os.path.join(args.out, "Loss_Trend_Full.png"),
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
nodes = torch.as_tensor(np.asarray(geometry["nodes"], dtype=np.float32)[:, :2], device="cuda")
my_data = model(nodes).cpu().numpy()
np.savez(os.path.join(args.out, "sol_pinn.npz"), u_pinn=my_data[:, 0] * data.norm_vel, v_pinn=my_data[:, 1] * data.norm_vel,
         p_pinn=my_data[:, 2] * data.norm_pre)
options.write_recap(os.path.join(args.out, "Test_Options.txt"), "Coronary_Flow_Steady", opt, with_initial_conditions=False)
total, train_vals, test_vals = pb.evaluate_all()
recap = {"engine": pb.plan.engine, "adam_seconds": t1 - t0, "bfgs_seconds": t2 - t1,
         "bfgs_iterations": pb.history["log"]["iter"][-1] - 100, "loss_global": total,
         "losses": {l.name: v for l, v in zip(pb.losses, train_vals)},
         "losses_test": {l.name: v for l, v in zip(pb.losses_test, test_vals)}}
with open(os.path.join(args.out, "Run_Summary.json"), "w") as fh:
    json.dump({"epochs": epochs, "n_pts": opt.n_pts, **recap}, fh, indent=2)
print(json.dumps(recap, indent=1))
