"""Throughput of the two optimiser rounds of the scripts (cavity_steady.py:246-247) on the BASELINE Cavity_Steady
workload: Adam on the device, SciPy BFGS driving the device step from the host (dense 2307 x 2307 inverse Hessian
in NumPy).  Development aid for the 'next' row of SURVEY.md 8(f)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pinns_fluid_dynamics_b200 as ns
from pinns_fluid_dynamics_b200 import loss_tables, problems
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
data = problems.build_baseline_config("Cavity_Steady", seed=1, PDE=n)
model = ns.TanhMLP(data.dim, data.hidden, data.out_dim, device="cuda", seed=3)
losses, ltest = loss_tables.build_loss_table(data)
pb = ns.OptimizationProblem(model.variables, losses, ltest)
ns.minimize(pb, "keras", ns.optimizers.Adam(learning_rate=1e-2), num_epochs=10)
torch.cuda.synchronize(); t0 = time.perf_counter()
ns.minimize(pb, "keras", ns.optimizers.Adam(learning_rate=1e-2), num_epochs=100)
torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"Adam round: {(t1 - t0) / 100 * 1e3:.3f} ms per epoch ({n} collocation points), loss {pb.evaluate()[0]:.4e}")
ns.minimize(pb, "scipy", "BFGS", num_epochs=2)      # warm-up: cuBLAS handle, float64 kernels
evals = [0]
orig = pb.evaluate_host
def counted(theta):
    evals[0] += 1
    return orig(theta)
pb.evaluate_host = counted
t0 = time.perf_counter()
ns.minimize(pb, "scipy", "BFGS", num_epochs=30)
t1 = time.perf_counter()
pb.evaluate_host = orig
print(f"BFGS round ({os.environ.get("PINN_BFGS", "device")} algebra): {(t1 - t0) / 30 * 1e3:.3f} ms per iteration, {evals[0] / 30:.2f} loss/gradient evaluations per iteration, "
      f"loss {pb.evaluate()[0]:.4e}")
