#!/usr/bin/env python
"""Poisson problem with mixed boundary conditions on the B200 framework: Examples/Poisson_Problem/poisson_misto.py, written
out with the facade's own names so the correspondence with the script is line by line (the other examples go through
``loss_tables``):

    -u_xx - u_yy = 2 sin x sin y  in (0, 2 pi)^2,   u = 0 on y = 0, 2 pi,   u_x = sin y on x = 0, 2 pi,   u = sin x sin y

    python examples/poisson_misto.py [--epochs N] [--out DIR]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pinns_fluid_dynamics_b200 as ns
from pinns_fluid_dynamics_b200 import residuals as operator          # the scripts' `tens_style as operator`
from pinns_fluid_dynamics_b200.residuals import PointSet

ap = argparse.ArgumentParser()
ap.add_argument("--epochs", type=int, default=7500)                  # poisson_misto.py:94
ap.add_argument("--out", default=os.path.dirname(os.path.abspath(__file__)))
args = ap.parse_args()

# %% Options (poisson_misto.py:21-36)
domain_W1 = domain_W2 = 2 * np.pi
u_exact = lambda x: np.sin(x[:, 0]) * np.sin(x[:, 1])
forcing = lambda x: 2 * np.sin(x[:, 0]) * np.sin(x[:, 1])
num_PDE, num_BC, num_test = 200, 20, 1000

# %% Initialisation (:38-60): the 2-20x3-1 tanh network and uniformly sampled point sets
rng = np.random.default_rng(1)
uniform = lambda n, lo, hi: np.asarray(lo, dtype=np.float64) + rng.random((n, 2)) * (np.asarray(hi, dtype=np.float64) - np.asarray(lo))
model = ns.TanhMLP(2, [20, 20, 20], 1, device="cuda", seed=1)
x_PDE = uniform(num_PDE, [0, 0], [domain_W1, domain_W2])
x_BC_x0, x_BC_x1 = uniform(num_BC, [0, 0], [0, domain_W2]), uniform(num_BC, [domain_W1, 0], [domain_W1, domain_W2])
x_BC_y0, x_BC_y1 = uniform(num_BC, [0, 0], [domain_W1, 0]), uniform(num_BC, [0, domain_W2], [domain_W1, domain_W2])
x_test = uniform(num_test, [0, 0], [domain_W1, domain_W2])
x_BC_D, x_BC_N = np.concatenate([x_BC_y0, x_BC_y1]), np.concatenate([x_BC_x0, x_BC_x1])
u_test, f, g = u_exact(x_test), forcing(x_PDE), np.sin(x_BC_N[:, 1])

# %% Residuals (:62-82): declarative forms instead of tape closures -- one fused kernel evaluates all terms of a point set
pde_set, bcd_set, bcn_set, test_set = PointSet(x_PDE, "PDE"), PointSet(x_BC_D, "BC_D"), PointSet(x_BC_N, "BC_N"), PointSet(x_test, "Test")
PDE = lambda: operator.poisson_pde(pde_set, f)                        # -laplacian(u) - f
BC_D = lambda: operator.dirichlet(bcd_set, 0, None)                   # u
BC_N = lambda: operator.normal_derivative(bcn_set, 0, 0, g)           # u_x - g

# %% Losses (:84-88)
losses = [ns.LossMeanSquares("PDE", PDE, weight=1e2),
          ns.LossMeanSquares("BC_D", BC_D),
          ns.LossMeanSquares("BC_N", BC_N)]
loss_test = ns.LossMeanSquares("fit", lambda: operator.dirichlet(test_set, 0, u_test))

# %% Training (:90-94)
pb = ns.OptimizationProblem(model.variables, losses, loss_test)
ns.minimize(pb, "keras", ns.optimizers.Adam(learning_rate=1e-2), num_epochs=10)
ns.minimize(pb, "scipy", "L-BFGS-B", num_epochs=args.epochs)

# %% Saving the loss history (:96-102) and the post-processing scatter data (:104-115)
history_file = os.path.join(args.out, "Poisson_Misto_history_loss.json")
pb.save_history(history_file)
history = ns.utils.load_json(history_file)
u_num = model(x_test).cpu().numpy()[:, 0]
np.savez(os.path.join(args.out, "Poisson_Misto_solution.npz"), x_test=x_test, u_exact=u_test, u_numerical=u_num)
print(f"engine {pb.plan.engine}: {len(history['log']['iter'])} history entries, final loss {history['log']['loss_global'][-1]:.3e}, "
      f"relative L2 error of u on the test points {np.linalg.norm(u_num - u_test) / np.linalg.norm(u_test):.3e}")
