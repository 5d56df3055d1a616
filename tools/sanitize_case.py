"""Small cases for compute-sanitizer (memcheck): every engine, ragged point counts, one step each."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pinns_fluid_dynamics_b200 as ns
from pinns_fluid_dynamics_b200 import loss_tables, problems
cases = [("cavity_unsteady", dict(PDE=333, BC=17, IC=9, Vel=3, Pres=1, Test=11, n_times=3, hidden=(128,) * 3)),
         ("cavity_unsteady", dict(PDE=97, BC=5, IC=5, Vel=1, Pres=1, Test=5, n_times=3, hidden=(64,) * 2)),
         ("cavity_steady", dict(PDE=211, BC=13, Vel=7, Pres=1, Test=9)),
         ("poiseuille_flow", dict(PDE=131, BC=21, Vel=5, Pres=0, Test=7))]
for name, kw in cases:
    data = problems.BUILDERS[name](seed=1, **kw)
    model = ns.TanhMLP(data.dim, data.hidden, data.out_dim, device="cuda", seed=3)
    losses, ltest = loss_tables.build_loss_table(data, faithful=False)
    pb = ns.OptimizationProblem(model.variables, losses, ltest)
    total, _, grad = pb.evaluate()
    pb.evaluate_all()
    y = model(torch.rand(77, data.dim, device="cuda"))
    torch.cuda.synchronize()
    print(name, pb.plan.engine, f"{total:.6e}", float(grad.norm()), tuple(y.shape))
