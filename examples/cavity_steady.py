#!/usr/bin/env python
"""Cavity_Steady on the B200 framework: Examples/Cavity_Steady/cavity_steady.py (options :37-58, grid + FEM fields :86-112,
point sets :114-153, loss table :204-231, training :236-247, saving :249-252, solution on the regular grid :254-278).

    python examples/cavity_steady.py [--epochs N] [--fem navier-stokes_cavity_steady.h5] [--pde N] [--out DIR]

``--fem`` names the FEniCS companion file of DataGeneration/fluid_solver_steady.py (VisualisationVector/0,1 on the 101 x 101
vertices); without it a smooth synthetic cavity field stands in for the fitting data (the reference data is not shipped).
"""
import argparse
import os

import numpy as np

from _common import HERE, read_or_write_options, train_and_save
import pinns_fluid_dynamics_b200 as ns
from pinns_fluid_dynamics_b200 import problems

ap = argparse.ArgumentParser()
ap.add_argument("--epochs", type=int, default=None, help="override TRAINING EPOCHS of the options file")
ap.add_argument("--pde", type=int, default=None, help="override POINTS PDE (BASELINE config 4: 1000000)")
ap.add_argument("--fem", default=None)
ap.add_argument("--out", default=os.path.join(HERE, "Test_Case_cavity_steady"))
args = ap.parse_args()

# Examples/Cavity_Steady/simulation_options.txt as checked in
opt = read_or_write_options("cavity_steady", dict(epochs=10000, noise_factor_bnd=0.01, noise_factor_fit=0.01),
                            {"PDE": 1000, "BC": 1000, "IC": 1000, "Vel": 100, "Pres": 1, "Test": 1000})
if args.pde is not None:
    opt.n_pts["PDE"] = args.pde
epochs = opt.epochs if args.epochs is None else args.epochs

fem = problems.load_fem_fields(args.fem) if args.fem else None
data = problems.cavity_steady(options=opt, seed=1, fem_fields=fem)
model = ns.TanhMLP(data.dim, data.hidden, data.out_dim, device="cuda", seed=1)
pb, recap = train_and_save("Cavity_Steady", data, model, opt, epochs, args.out)

# solution on the regular 100 x 100 grid, de-normalised (cavity_steady.py:254-278: what the contour plots show)
gx, gy = np.meshgrid(np.linspace(0, 1, 100), np.linspace(0, 1, 100))
grid = np.stack([gx.reshape(-1), gy.reshape(-1)], axis=-1)
y = model(grid).cpu().numpy()
u, v, p = (y[:, 0] * data.norm_vel).reshape(gx.shape), (y[:, 1] * data.norm_vel).reshape(gx.shape), (y[:, 2] * data.norm_pre).reshape(gx.shape)
np.savez(os.path.join(args.out, "Solution_Grid.npz"), grid_x=gx, grid_y=gy, u=u, v=v, p=p)
print(f"lid row mean u = {u[-1].mean():.4g} (lid velocity 500), wall rows |u| max = {np.abs(u[0]).max():.3g}")
