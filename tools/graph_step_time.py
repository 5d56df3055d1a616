"""Adam training step (loss step + update) per config: eager launches vs CUDA-graph replay.
usage: python tools/graph_step_time.py [steps]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pinns_fluid_dynamics_b200 as ns  # noqa: E402
from pinns_fluid_dynamics_b200 import loss_tables, problems  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
for name in ("Poiseuille_Flow", "Colliding_Flow", "Cavity_Steady", "Cavity_Unsteady"):
    row = []
    for flag in ("0", "1"):
        os.environ["PINN_CUDA_GRAPH"] = flag
        data = problems.build_baseline_config(name, seed=1)
        model = ns.TanhMLP(data.dim, data.hidden, data.out_dim, device="cuda", seed=2)
        losses, loss_test = loss_tables.build_loss_table(data)
        pb = ns.OptimizationProblem(model.variables, losses, loss_test)
        opt = ns.optimizers.Adam(learning_rate=1e-3)
        for _ in range(10):
            pb.training_step(opt)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            pb.training_step(opt)
        e1.record()
        torch.cuda.synchronize()
        row.append((e0.elapsed_time(e1) / steps, (time.perf_counter() - t0) * 1e3 / steps, pb._graph is not None))
    print(f"{name:16s} eager {row[0][0]:.4f} ms/step (wall {row[0][1]:.4f})   graph {row[1][0]:.4f} ms/step "
          f"(wall {row[1][1]:.4f}, replayed={row[1][2]})   x{row[0][0] / row[1][0]:.2f}")
