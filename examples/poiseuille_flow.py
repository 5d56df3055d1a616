#!/usr/bin/env python
"""Poiseuille_Flow on the B200 framework: Examples/Poiseuille_Flow/poiseuille_flow.py (options :37-58, channel geometry and
the analytic parabolic profile :60-100, point sets :102-164, loss table with the Neumann outflow terms :214-254, training
:259-266).  The analytic solution makes the result checkable: the script prints the velocity-profile error at mid-channel.

    python examples/poiseuille_flow.py [--epochs N] [--pde N] [--out DIR]
"""
import argparse
import os

import numpy as np

from _common import HERE, read_or_write_options, train_and_save
import pinns_fluid_dynamics_b200 as ns
from pinns_fluid_dynamics_b200 import problems

ap = argparse.ArgumentParser()
ap.add_argument("--epochs", type=int, default=None)
ap.add_argument("--pde", type=int, default=None, help="override POINTS PDE (BASELINE config 2: 10000)")
ap.add_argument("--out", default=os.path.join(HERE, "Test_Case_poiseuille_flow"))
args = ap.parse_args()

# Examples/Poiseuille_Flow/simulation_options.txt as checked in
opt = read_or_write_options("poiseuille_flow", dict(epochs=10000, noise_factor_bnd=0.0, noise_factor_fit=0.0),
                            {"PDE": 1000, "BC": 100, "IC": 100, "Vel": 10, "Pres": 0, "Test": 1000})
if args.pde is not None:
    opt.n_pts["PDE"] = args.pde
epochs = opt.epochs if args.epochs is None else args.epochs

data = problems.poiseuille_flow(options=opt, seed=1)
model = ns.TanhMLP(data.dim, data.hidden, data.out_dim, device="cuda", seed=1)
pb, recap = train_and_save("Poiseuille_Flow", data, model, opt, epochs, args.out)

# u(y) at mid-channel against the parabola u = (P_str - P_end) / (2 mu L) * y (2 delta - y)  (poiseuille_flow.py:72-80)
rho, mu, P_str, P_end, L, delta = 3100.0, 890.0, 1e6, 0.0, 1.0, 0.05
ys = np.linspace(0.0, 2 * delta, 41)
pts = np.stack([np.full_like(ys, 0.5 * L), ys], axis=-1)
u_pinn = model(pts).cpu().numpy()[:, 0] * data.norm_vel
u_ex = (P_str - P_end) / (2 * mu * L) * ys * (2 * delta - ys)
err = float(np.linalg.norm(u_pinn - u_ex) / np.linalg.norm(u_ex))
np.savez(os.path.join(args.out, "Profile_Mid_Channel.npz"), y=ys, u_pinn=u_pinn, u_exact=u_ex)
print(f"mid-channel velocity profile: relative L2 error {err:.3e} after {epochs} quasi-Newton epochs")
