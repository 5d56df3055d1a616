"""Per-kernel device times of one loss step of the 8x128 network (torch profiler, no ncu): development aid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pinns_fluid_dynamics_b200 as ns
from pinns_fluid_dynamics_b200 import loss_tables, problems
from torch.profiler import profile, ProfilerActivity
kw = dict(PDE=172800, BC=8, IC=8, Vel=1, Pres=1, Test=8, n_times=4, hidden=(128,) * 8)
data = problems.cavity_unsteady(seed=1, **kw)
model = ns.TanhMLP(data.dim, data.hidden, data.out_dim, device="cuda", seed=3)
losses, ltest = loss_tables.build_loss_table(data, faithful=True)
pb = ns.OptimizationProblem(model.variables, losses, ltest)
for _ in range(3): pb.plan.loss_and_grad(pb.flat)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): pb.plan.loss_and_grad(pb.flat)
    torch.cuda.synchronize()
agg = {}
for e in prof.events():
    if e.device_type.name == "CUDA" and e.name.startswith(("void pinn", "pinn", "tc_", "void tc_")) or "tc_" in e.name:
        k = e.name.split("(")[0][-40:]
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:8]:
    print(f"  {k:42s} n={a[0]:4d} avg={a[1]/a[0]:8.1f} us")
