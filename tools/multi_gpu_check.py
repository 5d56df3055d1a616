"""Run under torchrun on N GPUs: every rank evaluates its shard, one all-reduce joins them (the peer-memory kernel for the 3x32
network, NCCL for the 116 k-parameter 8x128 one), and rank 0 compares the global loss terms and gradient with the float64 Taylor
oracle of the WHOLE problem (both engines).  PINN_P2P_COMPARE=1 adds 310 graph-replayed Adam steps with each all-reduce path.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/multi_gpu_check.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import pinns_fluid_dynamics_b200 as ns
from pinns_fluid_dynamics_b200 import loss_tables, problems
from pinns_fluid_dynamics_b200.engine import assemble_losses, compile_problem
from oracle import reference_step, taylor

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cases = [("cavity_steady", dict(PDE=5001, BC=333, Vel=50, Pres=1, Test=100, noise_bnd=0.01, noise_fit=0.01), None),
         ("cavity_unsteady", dict(PDE=3001, BC=120, IC=77, Vel=3, Pres=1, Test=50, noise_bnd=0.05, noise_fit=0.05, n_times=3,
                                  hidden=(128,) * 8), None)]
for name, kw, _ in cases:
    data = problems.BUILDERS[name](seed=1, **kw)
    var = reference_step.glorot_uniform_variables(data.dim, data.hidden, data.out_dim, seed=11, bias_std=0.1)
    model = ns.TanhMLP(data.dim, data.hidden, data.out_dim, device="cuda")
    model.set_weights([v.numpy() for v in var])
    losses, ltest = loss_tables.build_loss_table(data)
    pb = ns.OptimizationProblem(model.variables, losses, ltest)
    total, values, grad = pb.evaluate()
    # oracle on the WHOLE problem: compile with world = 1
    cp = compile_problem([tuple(v.shape) for v in model.variables], losses, ltest, 0, 1)
    theta = torch.cat([v.reshape(-1) for v in var]).numpy()
    out = taylor.loss_and_grad(cp, theta)
    ref_total, ref_vals, _ = assemble_losses(cp, out[cp.n_params:])
    g = grad.double().cpu().numpy(); rg = out[:cp.n_params]
    worst = max(abs(v - rv) / abs(rv) for v, rv in zip(values, ref_vals) if rv != 0)
    if rank == 0:
        print(f"{name} [{pb.plan.engine}] world={world}: total rel err {abs(total-ref_total)/abs(ref_total):.2e}, worst term {worst:.2e}, "
              f"grad rel L2 {np.linalg.norm(g-rg)/np.linalg.norm(rg):.2e}", flush=True)

# ---- the two all-reduce paths side by side: NCCL in the graph vs the one-shot peer-memory kernel (PINN_P2P_ALLREDUCE=1) ----------
if os.environ.get("PINN_P2P_COMPARE", "0") == "1":
    import ctypes as C
    from pinns_fluid_dynamics_b200.api import Adam
    n_pde = int(os.environ.get("PINN_P2P_COMPARE_PDE", "200000"))
    finals, times = {}, {}
    for path in ("nccl", "p2p"):
        os.environ["PINN_P2P_ALLREDUCE"] = "1" if path == "p2p" else "0"
        data = problems.cavity_steady(seed=1, PDE=n_pde, BC=1000, Vel=100, Pres=1, Test=100, noise_bnd=0.01, noise_fit=0.01)
        model = ns.TanhMLP(data.dim, data.hidden, data.out_dim, device="cuda", seed=3)
        losses, ltest = loss_tables.build_loss_table(data)
        pb = ns.OptimizationProblem(model.variables, losses, ltest)
        opt = Adam(1e-3)
        for _ in range(10):
            pb.training_step(opt)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(300):
            pb.training_step(opt)
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 300], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times[path] = float(t.item())
        finals[path] = pb.flat.detach().double().cpu().numpy().copy()
        used = "p2p" if pb._p2p is not None else "nccl"
        timeouts = C.c_int32(0)
        if pb._p2p is not None:
            pb.plan.lib.pinn_p2p_status(pb._p2p, C.byref(timeouts))
        # every rank must hold the same parameters bit for bit
        mine = pb.flat.detach().clone(); ref = mine.clone(); dist.broadcast(ref, src=0)
        same = bool(torch.equal(mine, ref))
        if rank == 0 or not same:
            print(f"allreduce path requested {path}, used {used}: {times[path]*1e3:.1f} us per step ({n_pde} points in total), "
                  f"timeouts {timeouts.value}, ranks bit-identical: {same}", flush=True)
    if rank == 0:
        d = np.linalg.norm(finals["p2p"] - finals["nccl"]) / np.linalg.norm(finals["nccl"])
        print(f"p2p vs nccl after 310 Adam steps: relative difference of the parameters {d:.2e}; "
              f"step {times['nccl']*1e3:.1f} -> {times['p2p']*1e3:.1f} us world={world}", flush=True)
dist.destroy_process_group()
