"""Errors of the CUDA loss step against the float64 Taylor-mode oracle (development aid for precision experiments).
usage: python tools/accuracy_check.py [case] [n_pde]   (PINN_LIBPINNSTEP selects an alternative build)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pinns_fluid_dynamics_b200 as ns  # noqa: E402
from oracle import reference_step, taylor  # noqa: E402
from pinns_fluid_dynamics_b200 import loss_tables, problems  # noqa: E402
from pinns_fluid_dynamics_b200.engine import assemble_losses  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "cavity_steady"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
kw = dict(PDE=n, BC=200, Vel=50, Pres=1, Test=100, noise_bnd=0.01, noise_fit=0.01)
if case == "cavity_unsteady":
    kw.update(IC=200, n_times=4)
worst = [0.0, 0.0, 0.0]
for seed in (1, 2, 3):
    for bias_std in (0.0, 0.3):
        data = problems.BUILDERS[case](seed=seed, **kw)
        var = reference_step.glorot_uniform_variables(data.dim, data.hidden, data.out_dim, seed=seed + 10, bias_std=bias_std)
        model = ns.TanhMLP(data.dim, data.hidden, data.out_dim, device="cuda")
        model.set_weights([v.numpy() for v in var])
        losses, lt = loss_tables.build_loss_table(data)
        pb = ns.OptimizationProblem(model.variables, losses, lt)
        total, vals, grad = pb.evaluate()
        theta = torch.cat([v.reshape(-1) for v in var]).numpy()
        out = taylor.loss_and_grad(pb.compiled, theta)
        rt, rv, _ = assemble_losses(pb.compiled, out[pb.compiled.n_params:])
        rg = out[:pb.compiled.n_params]
        e_loss = abs(total - rt) / abs(rt)
        e_term = max(abs(v - r) / abs(r) for v, r in zip(vals[:3], rv[:3]) if r != 0)
        e_grad = np.linalg.norm(grad.double().cpu().numpy() - rg) / np.linalg.norm(rg)
        worst = [max(worst[0], e_loss), max(worst[1], e_term), max(worst[2], e_grad)]
        print(f"seed {seed} bias_std {bias_std}: loss {e_loss:.2e}  worst PDE term {e_term:.2e}  grad {e_grad:.2e}")
print(f"{case} n={n} [{pb.plan.engine}] worst: loss {worst[0]:.2e} (tol 1e-5)  PDE term {worst[1]:.2e} (tol 1e-5)  grad {worst[2]:.2e} (tol 1e-4)")
