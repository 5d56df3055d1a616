#!/usr/bin/env python
"""Cavity_Unsteady on the B200 framework: Examples/Cavity_Unsteady/cavity_unsteady.py (options :37-58, space-time grid :86-101,
per-time-step FEM files :103-113, point sets incl. the initial condition :125-163, loss table with the d/dt momentum terms
:178-199 and :223-250, training :255-266).

    python examples/cavity_unsteady.py [--epochs N] [--fem-folder DIR] [--n-times N] [--hidden 8x128] [--pde N] [--out DIR]

``--fem-folder`` holds navier-stokes_SI_cavity_unsteady_{step:05d}.h5 of DataGeneration/fluid_solver_unsteady.py (one file per
time step, pressure mean subtracted per step); without it a synthetic impulsively-started-lid field stands in.  ``--hidden 8x128``
selects the wide network of BASELINE config 5 (tensor-core engine); the script's own network is 3-32x3-3.
"""
import argparse
import os

import numpy as np

from _common import HERE, read_or_write_options, train_and_save
import pinns_fluid_dynamics_b200 as ns
from pinns_fluid_dynamics_b200 import problems

ap = argparse.ArgumentParser()
ap.add_argument("--epochs", type=int, default=None)
ap.add_argument("--pde", type=int, default=None)
ap.add_argument("--fem-folder", default=None)
ap.add_argument("--n-times", type=int, default=None, help="time steps of the space-time grid (script: T / dt = 100)")
ap.add_argument("--hidden", default="3x32", help="LAYERSxWIDTH, e.g. 3x32 (script) or 8x128 (BASELINE config 5)")
ap.add_argument("--out", default=os.path.join(HERE, "Test_Case_cavity_unsteady"))
args = ap.parse_args()

# Examples/Cavity_Unsteady/simulation_options.txt as checked in
opt = read_or_write_options("cavity_unsteady", dict(epochs=10000, noise_factor_bnd=0.0, noise_factor_fit=0.0),
                            {"PDE": 1000, "BC": 100, "IC": 100, "Vel": 100, "Pres": 1, "Test": 1000})
if args.pde is not None:
    opt.n_pts["PDE"] = args.pde
epochs = opt.epochs if args.epochs is None else args.epochs
layers, width = (int(t) for t in args.hidden.split("x"))

data = problems.cavity_unsteady(options=opt, seed=1, hidden=(width,) * layers, n_times=args.n_times, fem_fields=args.fem_folder)
model = ns.TanhMLP(data.dim, data.hidden, data.out_dim, device="cuda", seed=1)
pb, recap = train_and_save("Cavity_Unsteady", data, model, opt, epochs, args.out)

# the fields at the first, middle and last time of the horizon on the regular grid, de-normalised (cavity_unsteady.py:281-330)
T = float(data.consts["T"])
gx, gy = np.meshgrid(np.linspace(0, 1, 100), np.linspace(0, 1, 100))
frames = {}
for tag, t in (("t0", 0.0), ("tmid", 0.5 * T), ("tend", T)):
    pts = np.stack([np.full(gx.size, t), gx.reshape(-1), gy.reshape(-1)], axis=-1)
    y = model(pts).cpu().numpy()
    frames[f"u_{tag}"] = (y[:, 0] * data.norm_vel).reshape(gx.shape)
    frames[f"v_{tag}"] = (y[:, 1] * data.norm_vel).reshape(gx.shape)
    frames[f"p_{tag}"] = (y[:, 2] * data.norm_pre).reshape(gx.shape)
np.savez(os.path.join(args.out, "Solution_Grid.npz"), grid_x=gx, grid_y=gy, **frames)
