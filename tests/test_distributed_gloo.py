"""N > 1 host path on CPU: two gloo ranks shard every point set, evaluate their shards (oracle engine
injected), all-reduce the [P+T] vector and must reproduce the single-rank result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _build(engine_factory, case="cavity_unsteady"):
    import pinns_fluid_dynamics_b200 as ns
    from pinns_fluid_dynamics_b200 import loss_tables, problems
    if case == "cavity_unsteady":
        data = problems.cavity_unsteady(seed=3, PDE=101, BC=9, IC=5, Vel=3, Pres=1, Test=7, noise_bnd=0.05,
                                        noise_fit=0.05, n_times=3)
    else:   # ns.Loss over |mean p|: its point set must stay whole on rank 0
        data = problems.colliding_flow_pressmean(seed=3, num_pde=101, num_bc=9, num_test=7, num_pres=11)
    model = ns.TanhMLP(data.dim, data.hidden, data.out_dim, device="cpu", seed=8)
    losses, ltest = loss_tables.build_loss_table(data)
    return ns, ns.OptimizationProblem(model.variables, losses, ltest, engine_factory=engine_factory)


def _worker(rank, world, port, q, case="cavity_unsteady"):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.taylor import TaylorEngine
    ns, pb = _build(TaylorEngine, case)
    assert pb.world == world and pb.rank == rank
    n_local = {cs.pointset.name: cs.n_local for cs in pb.compiled.sets}
    total, vals, grad = pb.evaluate()
    t2, tr, te = pb.evaluate_all()
    ns.minimize(pb, "keras", ns.optimizers.Adam(learning_rate=1e-2), num_epochs=3)
    q.put((rank, total, vals, grad.numpy(), te, n_local, pb.flat.numpy().copy()))
    dist.destroy_process_group()


@pytest.mark.parametrize("case", ["cavity_unsteady", "colliding_flow_pressmean"])
def test_two_ranks_reproduce_single_rank(case):
    from oracle.taylor import TaylorEngine
    ns, pb1 = _build(TaylorEngine, case)
    total1, vals1, grad1 = pb1.evaluate()
    _, _, te1 = pb1.evaluate_all()
    ns.minimize(pb1, "keras", ns.optimizers.Adam(learning_rate=1e-2), num_epochs=3)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, case)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(2)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, total, vals, grad, te, n_local, flat in res:
        assert abs(total - total1) <= 1e-12 * abs(total1)
        assert np.allclose(vals, vals1, rtol=1e-12) and np.allclose(te, te1, rtol=1e-12)
        assert np.linalg.norm(grad - grad1.numpy()) <= 1e-12 * np.linalg.norm(grad1.numpy())
        assert np.allclose(flat, pb1.flat.numpy(), rtol=0, atol=1e-7)        # identical updates on every rank
    # contiguous shards: first rank takes the extra row; the 1-point pressure set lives on rank 0 only
    assert res[0][5]["PDE"] == 51 and res[1][5]["PDE"] == 50
    if case == "cavity_unsteady":
        assert res[0][5]["Pres"] == 1 and res[1][5]["Pres"] == 0
    else:   # the |mean| term's set is not sharded
        assert res[0][5]["pres"] == 11 and res[1][5]["pres"] == 0
