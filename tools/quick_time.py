"""Quick device timing of the loss step (development aid; bench.py is the contract)."""
import ctypes as C
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pinns_fluid_dynamics_b200 as ns
from pinns_fluid_dynamics_b200 import loss_tables, problems, _capi

name = sys.argv[1] if len(sys.argv) > 1 else "cavity_steady"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
kw = dict(PDE=n, BC=1000, Vel=100, Pres=1, Test=1000, noise_bnd=0.01, noise_fit=0.01)
if name == "cavity_unsteady":
    kw.update(IC=1000, n_times=4)
if len(sys.argv) > 3:
    h, l = sys.argv[3].split("x")
    kw.update(hidden=(int(h),) * int(l))
data = problems.BUILDERS[name](seed=1, **kw)
model = ns.TanhMLP(data.dim, data.hidden, data.out_dim, device="cuda", seed=3)
losses, ltest = loss_tables.build_loss_table(data, faithful=(name.startswith("cavity")))
pb = ns.OptimizationProblem(model.variables, losses, ltest)
lib = _capi.load()
lib.pinn_plan_enable_timing(pb.plan.handle, 1)
for _ in range(3):
    pb.plan.loss_and_grad(pb.flat)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 20 if pb.plan.engine.startswith('fused') else 3
e0.record()
for _ in range(K):
    pb.plan.loss_and_grad(pb.flat)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
kms = C.c_float()
if pb.plan.engine.startswith("fused"):
    lib.pinn_plan_kernel_time_ms(pb.plan.handle, 2, C.byref(kms))
else:
    kms.value = ms
d, H, L, O = pb.compiled.mlp
Cc = 3 + d
F = 3 * (2 * d * H + (L - 1) * Cc * 2 * H * H + Cc * 2 * H * O)
print(f"{name} [{pb.plan.engine}] n={n}: step {ms:.3f} ms  ({n/ms*1e3:.3e} pts/s); collocation kernel {kms.value:.3f} ms -> "
      f"{n*F/kms.value*1e-9:.2f} TFLOP/s algorithmic; launches/step {pb.plan.last_launch_count()}")
