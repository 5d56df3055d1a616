"""Accuracy of the wide-network engines against the float64 Taylor oracle (development aid):
per-term relative error of the loss values, relative error of the total, relative L2 error of the gradient."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pinns_fluid_dynamics_b200 as ns
from pinns_fluid_dynamics_b200 import loss_tables, problems
from pinns_fluid_dynamics_b200.engine import assemble_losses
from oracle import reference_step, taylor

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
for seed in (1, 2, 3):
    kw = dict(PDE=n, BC=100, IC=100, Vel=3, Pres=1, Test=50, noise_bnd=0.05, noise_fit=0.05, n_times=3, hidden=(128,) * 8)
    data = problems.cavity_unsteady(seed=seed, **kw)
    var = reference_step.glorot_uniform_variables(data.dim, data.hidden, data.out_dim, seed=seed + 10, bias_std=0.1)
    model = ns.TanhMLP(data.dim, data.hidden, data.out_dim, device="cuda")
    model.set_weights([v.numpy() for v in var])
    losses, ltest = loss_tables.build_loss_table(data, faithful=True)
    pb = ns.OptimizationProblem(model.variables, losses, ltest)
    total, values, grad = pb.evaluate()
    theta = torch.cat([v.reshape(-1) for v in var]).numpy()
    out = taylor.loss_and_grad(pb.compiled, theta)
    ref_total, ref_vals, _ = assemble_losses(pb.compiled, out[pb.compiled.n_params:])
    g = grad.double().cpu().numpy(); rg = out[:pb.compiled.n_params]
    terr = max(abs(v - rv) / abs(rv) for v, rv in zip(values, ref_vals) if rv != 0)
    worst = max(((abs(v - rv) / abs(rv), l.name) for l, v, rv in zip(pb.losses, values, ref_vals) if rv != 0))
    print(f"[{pb.plan.engine}] seed {seed} n={n}: total rel err {abs(total-ref_total)/abs(ref_total):.2e}, worst term {worst[0]:.2e} ({worst[1]}), "
          f"grad rel L2 {np.linalg.norm(g-rg)/np.linalg.norm(rg):.2e}")
