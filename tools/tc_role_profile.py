"""Where do the warp roles of tc_layer wait?  Needs a -DPINN_TC_PROFILE build of the library
(PINN_LIBPINNSTEP=...): CTA 0 accumulates clock64 intervals per role over every tc_layer launch of one step."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pinns_fluid_dynamics_b200 as ns
from pinns_fluid_dynamics_b200 import loss_tables, problems, _capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 172800
kw = dict(PDE=n, BC=8, IC=8, Vel=1, Pres=1, Test=8, n_times=4, hidden=(128,) * 8)
data = problems.cavity_unsteady(seed=1, **kw)
model = ns.TanhMLP(data.dim, data.hidden, data.out_dim, device="cuda", seed=3)
losses, ltest = loss_tables.build_loss_table(data, faithful=True)
pb = ns.OptimizationProblem(model.variables, losses, ltest)
lib = _capi.load()
buf = (C.c_ulonglong * 32)()
for _ in range(2):
    pb.plan.loss_and_grad(pb.flat)
lib.pinn_debug_tc_prof(buf, 1)
pb.plan.loss_and_grad(pb.flat)
lib.pinn_debug_tc_prof(buf, 1)
names = ["producer: wait empty slot", "producer: whole loop", "mma: wait tmem buffer drained", "mma: wait stage full", "mma: whole loop",
         "epilogue: wait chunk (tfull)", "epilogue: drain chunk", "epilogue: elementwise + global"]
tot = buf[4]
print(f"CTA 0, all tc_layer launches of one step ({n} collocation points + small sets); cycles, share of the MMA warp's loop")
for i, nm in enumerate(names):
    print(f"  {nm:34s} {buf[i]:12d}  {buf[i] / tot * 100:6.1f} %")
