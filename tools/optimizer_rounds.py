"""Throughput of the two optimiser rounds of the scripts (cavity_steady.py:246-247) on the BASELINE Cavity_Steady
workload: Adam on the device, SciPy BFGS driving the device step from the host (dense 2307 x 2307 inverse Hessian
in NumPy).  Development aid for the 'next' row of SURVEY.md 8(f)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pinns_fluid_dynamics_b200 as ns
from pinns_fluid_dynamics_b200 import loss_tables, problems
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
N_BFGS = int(sys.argv[2]) if len(sys.argv) > 2 else 300      # the scripts run 10 000: per-round set-up must not count
data = problems.build_baseline_config("Cavity_Steady", seed=1, PDE=n)
model = ns.TanhMLP(data.dim, data.hidden, data.out_dim, device="cuda", seed=3)
losses, ltest = loss_tables.build_loss_table(data)
pb = ns.OptimizationProblem(model.variables, losses, ltest)
ns.minimize(pb, "keras", ns.optimizers.Adam(learning_rate=1e-2), num_epochs=10)
torch.cuda.synchronize(); t0 = time.perf_counter()
ns.minimize(pb, "keras", ns.optimizers.Adam(learning_rate=1e-2), num_epochs=100)
torch.cuda.synchronize(); t1 = time.perf_counter()
adam_ms = (t1 - t0) / 100 * 1e3
print(f"Adam round: {(t1 - t0) / 100 * 1e3:.3f} ms per epoch ({n} collocation points), loss {pb.evaluate()[0]:.4e}")
ns.minimize(pb, "scipy", "BFGS", num_epochs=4)      # warm-up: pinned buffers, first graph capture, allocator
evals = [0]
orig = pb.evaluate_host
def counted(theta):
    evals[0] += 1
    return orig(theta)
pb.evaluate_host = counted
t0 = time.perf_counter()
ns.minimize(pb, "scipy", "BFGS", num_epochs=N_BFGS + 1)
t1 = time.perf_counter()
pb.evaluate_host = orig
nfev = evals[0] or int(pb.last_result.nfev)
nit = int(pb.last_result.nit)
print(f"BFGS round ({os.environ.get('PINN_BFGS', 'device')} algebra): {(t1 - t0) / nit * 1e3:.3f} ms per iteration over {nit} iterations, {nfev / nit:.2f} loss/gradient evaluations per iteration, "
      f"loss {pb.evaluate()[0]:.4e}")
sec = getattr(pb.last_result, "seconds", None)
if sec:
    print(f"  wall time inside the round: {sec['n_eval']} evaluations {sec['eval'] / max(sec['n_eval'], 1) * 1e3:.3f} ms each, "
          f"{sec['n_accept']} accept+update+direction {sec['accept'] / max(sec['n_accept'], 1) * 1e3:.3f} ms each, "
          f"rest (line-search logic, history) {((t1 - t0) - sec['eval'] - sec['accept']) / nit * 1e3:.3f} ms per iteration")
    print(f"  iteration / (evaluations per iteration x one evaluation) = {((t1 - t0) / nit) / (nfev / nit * sec['eval'] / max(sec['n_eval'], 1)):.2f}")
