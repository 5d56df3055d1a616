"""nisaba-shaped facade: the names the reference's example scripts call, backed by libpinnstep.so.

    ns.LossMeanSquares(name, eval_roots, weight=1.0, normalization=1.0)   cavity_steady.py:204,212-225
    ns.Loss(name, eval_loss, normalization, weight, non_negative)          colliding_flow_pressmean.py:196
    ns.OptimizationProblem(variables, losses, losses_test, callbacks=[])   cavity_steady.py:242
    ns.minimize(pb, 'keras', optimizer, num_epochs)                        cavity_steady.py:246
    ns.minimize(pb, 'scipy', 'BFGS' | 'L-BFGS-B', num_epochs)              cavity_steady.py:247, poisson.py:75
    ns.utils.HistoryPlotCallback / load_json / plot_history, pb.save_history   cavity_steady.py:243-245, poisson.py:81-83
    ns.config.get_dtype()                                                   poisson.py:47

``eval_roots`` is a declarative ``ResidualForm`` (or a zero-argument callable returning one, so the
scripts' ``lambda: PDE_MOM(0)`` idiom still reads the same) instead of a TensorFlow closure.
History bookkeeping follows the saved History_Loss.json files: entries every 10
iterations of each round plus iteration 0 of the round, ``loss_global = sum_t weight_t*log_t``,
test losses excluded from the total, rounds named ``keras_<Optimizer>`` / ``scipy_<method>``.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import math
from typing import Callable, List, Optional, Sequence, Union

import numpy as np
import torch

from . import _capi
from .engine import CompiledProblem, CudaPlan, assemble_losses, compile_problem
from .residuals import ResidualForm

LOG_STRIDE = 10


class config:
    @staticmethod
    def get_dtype():
        """The reference computes in float64 (Model.json); this build computes in FP32 by design."""
        return torch.float32


# ------------------------------------------------------------------------------------------------
# model
# ------------------------------------------------------------------------------------------------

class TanhMLP:
    """tf.keras.Sequential([Dense(H, tanh)] * L + [Dense(O)]) with GlorotUniform / zeros init
    (cavity_steady.py:205-210, Model.json).  ``variables`` are views, in Keras order and layout, of
    one flat FP32 vector -- the vector the kernels and optimisers work on."""

    def __init__(self, dim: int, hidden: Sequence[int], out_dim: int, device=None, seed: Optional[int] = None):
        if len(set(hidden)) != 1:
            raise ValueError("all hidden layers must have the same width")
        self.dim, self.hidden, self.out_dim = int(dim), [int(h) for h in hidden], int(out_dim)
        sizes = [self.dim] + self.hidden + [self.out_dim]
        self.shapes = []
        for i in range(len(sizes) - 1):
            self.shapes += [(sizes[i], sizes[i + 1]), (sizes[i + 1],)]
        n = sum(int(np.prod(s)) for s in self.shapes)
        if device is None:
            device = "cuda" if torch.cuda.is_available() else "cpu"
        self.flat = torch.zeros(n, dtype=torch.float32, device=device)
        self.variables: List[torch.Tensor] = []
        off = 0
        for s in self.shapes:
            k = int(np.prod(s))
            self.variables.append(self.flat[off:off + k].view(*s))
            off += k
        rng = np.random.default_rng(seed)
        w = []
        for i in range(len(sizes) - 1):
            lim = math.sqrt(6.0 / (sizes[i] + sizes[i + 1]))
            w += [rng.uniform(-lim, lim, size=(sizes[i], sizes[i + 1])), np.zeros(sizes[i + 1])]
        self.set_weights(w)

    def set_weights(self, arrays) -> None:
        if len(arrays) != len(self.variables):
            raise ValueError("wrong number of weight arrays")
        for v, a in zip(self.variables, arrays):
            a = torch.as_tensor(np.asarray(a), dtype=torch.float32)
            if tuple(a.shape) != tuple(v.shape):
                raise ValueError(f"weight shape {tuple(a.shape)} does not match {tuple(v.shape)}")
            v.copy_(a)

    def get_weights(self) -> List[np.ndarray]:
        return [v.detach().cpu().numpy().copy() for v in self.variables]

    def __call__(self, x) -> torch.Tensor:
        """model(x): values only, [n, O] (post-processing grids, cavity_steady.py:254-268)."""
        import ctypes as C
        if not self.flat.is_cuda:
            raise _capi.PinnLibraryError("model(x) runs on the GPU only (no CPU fallback)")
        lib = _capi.load()
        x = torch.as_tensor(x, dtype=torch.float32, device=self.flat.device).contiguous()
        if x.dim() != 2 or x.shape[1] != self.dim:
            raise ValueError(f"model(x) takes [n, {self.dim}] points, got {tuple(x.shape)}")
        y = torch.empty(x.shape[0], self.out_dim, dtype=torch.float32, device=self.flat.device)
        mlp = _capi.MlpDesc(self.dim, self.hidden[0], len(self.hidden), self.out_dim)
        stream = C.c_void_p(torch.cuda.current_stream(self.flat.device).cuda_stream)
        _capi.check(lib.pinn_forward(C.byref(mlp), C.c_void_p(self.flat.data_ptr()), C.c_void_p(x.data_ptr()),
                                     x.shape[0], C.c_void_p(y.data_ptr()), self.flat.device.index or 0, stream),
                    "pinn_forward")
        return y

    def to_json(self) -> str:
        """Architecture record in the shape of Keras' Model.json (cavity_steady.py:250-251)."""
        layers = [{"class_name": "InputLayer", "config": {"batch_input_shape": [None, self.dim], "dtype": "float32"}}]
        for i, h in enumerate(self.hidden + [self.out_dim]):
            act = "tanh" if i < len(self.hidden) else "linear"
            layers.append({"class_name": "Dense", "config": {
                "name": f"dense_{i}", "units": h, "activation": act, "use_bias": True, "dtype": "float32",
                "kernel_initializer": {"class_name": "GlorotUniform", "config": {"seed": None}},
                "bias_initializer": {"class_name": "Zeros", "config": {}}}})
        return json.dumps({"class_name": "Sequential", "config": {"name": "sequential", "layers": layers},
                           "backend": "pinns_fluid_dynamics_b200"})

    def save_weights(self, path: str) -> None:
        """``model.save_weights(f"{saving_folder}/Weights.h5")`` (cavity_steady.py:252).  ``*.h5`` / ``*.hdf5``: an HDF5 file
        with Keras' dataset paths (``dense/dense/kernel:0``, ``dense_1/dense_1/bias:0``, ...; float64 like the reference's
        files) written by ``h5lite`` -- the reference's tooling and ``load_weights`` read it back.  Anything else: ``.npz``
        with the same dataset paths as keys."""
        from . import h5lite
        arrays = self.get_weights()
        if path.endswith((".h5", ".hdf5")):
            h5lite.save_keras_dense_weights(path, arrays)
            return
        names = {}
        for i in range(len(arrays) // 2):
            layer = "dense" if i == 0 else f"dense_{i}"
            names[f"{layer}/{layer}/kernel:0"] = arrays[2 * i]
            names[f"{layer}/{layer}/bias:0"] = arrays[2 * i + 1]
        np.savez(path, **names)

    def load_weights(self, path: str) -> None:
        """``model.load_weights(path)``: a Keras ``Weights.h5`` (the reference's Test_Case folders, or ``save_weights``) or
        the ``.npz`` of ``save_weights``.  Layers are taken in the order of their Keras index (dense, dense_1, ...)."""
        from . import h5lite
        if path.endswith((".h5", ".hdf5")):
            self.set_weights(h5lite.load_keras_dense_weights(path))
            return
        with np.load(path if path.endswith(".npz") else path + ".npz") as z:
            layers = {}
            for key in z.files:
                parts = key.split("/")
                layers.setdefault(parts[0], {})[parts[-1]] = z[key]

        def index(name: str) -> int:
            tail = name.rsplit("_", 1)
            return int(tail[1]) if len(tail) == 2 and tail[1].isdigit() else 0
        arrays = []
        for name in sorted(layers, key=index):
            arrays += [layers[name]["kernel:0"], layers[name]["bias:0"]]
        self.set_weights(arrays)


# ------------------------------------------------------------------------------------------------
# losses
# ------------------------------------------------------------------------------------------------

class LossMeanSquares:
    """value = mean(roots**2) / normalization; history fields weight / non_negative / display_sqrt."""

    def __init__(self, name: str, eval_roots: Union[ResidualForm, Callable[[], ResidualForm]], *,
                 weight: float = 1.0, normalization: float = 1.0):
        form = eval_roots() if callable(eval_roots) and not isinstance(eval_roots, ResidualForm) else eval_roots
        if not isinstance(form, ResidualForm):
            raise TypeError("eval_roots must be a ResidualForm (see pinns_fluid_dynamics_b200.residuals) "
                            "or a callable returning one; arbitrary closures cannot be fused into the kernel")
        self.name, self.form = name, form
        self.weight, self.normalization = float(weight), float(normalization)
        self.non_negative, self.display_sqrt = True, True


class Loss:
    """``ns.Loss(name, eval_loss, normalization, weight, non_negative)`` -- a scalar loss, value = eval_loss() /
    normalization.  The reference's only use is ``ns.Loss('PRESS_0', lambda: PRESS_0(x_pres), normalization=1e0,
    weight=1e-2, non_negative=True)`` with PRESS_0 = |mean(model(x)[:, 2])| (colliding_flow_pressmean.py:176-179,196);
    that scalar is served on the fused path as the ``abs_mean`` reduction of a residual form
    (``residuals.mean_value``).  Any other scalar closure is rejected loudly."""

    def __init__(self, name, eval_loss, *, weight=1.0, normalization=1.0, non_negative=False):
        form = eval_loss() if callable(eval_loss) and not isinstance(eval_loss, ResidualForm) else eval_loss
        if not isinstance(form, ResidualForm) or form.reduction != "abs_mean":
            raise NotImplementedError(
                "ns.Loss serves |mean(residual)| forms (residuals.mean_value, colliding_flow_pressmean.py:196); "
                "an arbitrary scalar closure cannot be fused into the kernel")
        self.name, self.form = name, form
        self.weight, self.normalization = float(weight), float(normalization)
        self.non_negative, self.display_sqrt = bool(non_negative), False


# ------------------------------------------------------------------------------------------------
# problem
# ------------------------------------------------------------------------------------------------

def out_of_env(name: str) -> bool:
    """True when the environment switches feature ``name`` off (NAME=0)."""
    return os.environ.get(name, "1") == "0"


class OptimizationProblem:
    def __init__(self, variables, losses, losses_test=None, callbacks=None, *, engine_factory=None,
                 process_group=None):
        self.variables = list(variables)
        self.losses = list(losses)
        if losses_test is None:
            losses_test = []
        elif isinstance(losses_test, LossMeanSquares):
            losses_test = [losses_test]       # poisson.py:69,72 passes a single loss
        self.losses_test = list(losses_test)
        self.callbacks = [] if callbacks is None else callbacks
        self.flat = self._flat_view(self.variables)
        self.group = process_group
        import torch.distributed as dist
        self._dist = dist if (dist.is_available() and dist.is_initialized()) else None
        self.rank = self._dist.get_rank(self.group) if self._dist else 0
        self.world = self._dist.get_world_size(self.group) if self._dist else 1
        self.compiled: CompiledProblem = compile_problem([tuple(v.shape) for v in self.variables],
                                                         self.losses, self.losses_test, self.rank, self.world)
        # the plan (device copies of the point sets, workspace) lives on the device of the parameters
        self.plan = engine_factory(self.compiled) if engine_factory else CudaPlan(self.compiled, device=self.flat.device)
        self._graph, self._graph_opt, self._graph_sumsq, self._eager_steps = None, None, None, 0
        self._comm, self._comm_tried = None, False
        self._h_theta = self._h_out = self._perm_np = None      # pinned staging of evaluate_host
        self.iteration = 0
        self.history = {
            "log": {"iter": [], "round": [], "iter_round": [], "loss_global": []},
            "losses": {l.name: {"weight": l.weight, "non_negative": l.non_negative,
                                "display_sqrt": l.display_sqrt, "log": []} for l in self.losses},
            "losses_test": {l.name: {"weight": l.weight, "non_negative": l.non_negative,
                                     "display_sqrt": l.display_sqrt, "log": []} for l in self.losses_test},
            "log_rounds": {"rounds": [], "iteration_start": []},
        }

    @staticmethod
    def _flat_view(variables) -> torch.Tensor:
        base = variables[0]._base if variables[0]._base is not None else None
        n = sum(v.numel() for v in variables)
        if base is not None and base.dim() == 1 and base.numel() == n and base.is_contiguous():
            off, ok = 0, True
            for v in variables:
                ok &= v._base is base and v.storage_offset() == off and v.is_contiguous()
                off += v.numel()
            if ok:
                return base
        raise ValueError("variables must be the views of one flat FP32 vector in Keras order "
                         "(use TanhMLP(...).variables)")

    # ---- evaluation ---------------------------------------------------------------------------
    def _reduce(self, out: torch.Tensor) -> torch.Tensor:
        """SUM of the [P+T] vector over the ranks: the only cross-GPU dependency of a step.  On CUDA it goes through the
        library's own NCCL communicator (``pinn_allreduce_sum``: a plain ncclAllReduce on the current stream, which a CUDA
        graph captures as one node); ``torch.distributed`` serves CPU / gloo groups and ``PINN_OWN_NCCL=0``."""
        if self._dist is not None and self.world > 1:
            p2p = self._p2p_ctx() if out.is_cuda else None
            if p2p is not None and out.numel() <= self._p2p_cap:
                stream = C.c_void_p(torch.cuda.current_stream(out.device).cuda_stream)
                _capi.check(self.plan.lib.pinn_p2p_allreduce_sum(p2p, C.c_void_p(out.data_ptr()), out.numel(), stream), "pinn_p2p_allreduce_sum")
                return out
            comm = self._own_comm() if out.is_cuda else None
            if comm is not None:
                stream = C.c_void_p(torch.cuda.current_stream(out.device).cuda_stream)
                _capi.check(self.plan.lib.pinn_allreduce_sum(comm, C.c_void_p(out.data_ptr()), out.numel(), stream), "pinn_allreduce_sum")
            else:
                self._dist.all_reduce(out, op=self._dist.ReduceOp.SUM, group=self.group)
        return out

    _p2p = None
    _p2p_tried = False
    _p2p_cap = 0

    def _p2p_ctx(self):
        """One-shot all-reduce over NVLink peer memory (``pinn_p2p_*``, csrc/p2p.cuh) for the ranks of ONE node: every rank's
        receive block is mapped into its peers through CUDA IPC; the 64-byte handles travel over torch.distributed once.
        Used for the default group of up to 8 ranks on one host when every rank could map every peer; otherwise None and the
        NCCL path serves (also with ``PINN_P2P_ALLREDUCE=0``).  Measured on 8 B200s, 1 M points per rank: 716 -> 676 us per
        step (profiles/scaling_r02.md)."""
        if self._p2p_tried:
            return self._p2p
        self._p2p_tried = True
        if (out_of_env("PINN_P2P_ALLREDUCE") or self.group is not None or not isinstance(self.plan, CudaPlan) or self.world > 8
                or int(self.plan.out.numel()) > 16384):      # one CTA per rank: latency-bound vectors only (8x128 has 116 k parameters)
            return None
        import socket
        lib, dev = self.plan.lib, self.flat.device
        index = dev.index if dev.index is not None else torch.cuda.current_device()
        cap = int(self.plan.out.numel())
        handle = (C.c_ubyte * 64)()
        ctx = C.c_void_p()
        ok = lib.pinn_p2p_create(self.world, self.rank, index, cap, handle, C.byref(ctx)) == 0
        # one node only: the host name must agree on all ranks
        names = [None] * self.world
        self._dist.all_gather_object(names, socket.gethostname())
        ok = ok and len(set(names)) == 1
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
        every = torch.empty(64 * self.world, dtype=torch.uint8, device=dev)
        self._dist.all_gather_into_tensor(every, mine)
        if ok:
            raw = (C.c_ubyte * (64 * self.world))(*every.cpu().tolist())
            ok = lib.pinn_p2p_connect(ctx, raw) == 0
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        self._dist.all_reduce(flag, op=self._dist.ReduceOp.MIN)          # all ranks or none
        if int(flag.item()) == 1:
            self._p2p, self._p2p_cap = ctx, cap
        elif ctx.value:
            lib.pinn_p2p_destroy(ctx)
        return self._p2p

    def allreduce_timeouts(self) -> int:
        """Waits of the peer-memory all-reduce that timed out (a peer died or never launched): 0 in a healthy run."""
        if self._p2p is None:
            return 0
        n = C.c_int32(0)
        _capi.check(self.plan.lib.pinn_p2p_status(self._p2p, C.byref(n)), "pinn_p2p_status")
        return int(n.value)

    def close(self) -> None:
        """Release the peer-memory mappings (every rank, after the last step; the ranks should pass a barrier first)."""
        if self._p2p is not None:
            ctx, self._p2p = self._p2p, None
            self._graph = None
            self.plan.lib.pinn_p2p_destroy(ctx)

    def _own_comm(self):
        """NCCL communicator of the default group created through the C ABI (the 128-byte unique id travels over
        torch.distributed once).  None when it cannot be used (sub-groups, non-CUDA plans, PINN_OWN_NCCL=0)."""
        if self._comm_tried:
            return self._comm
        self._comm_tried = True
        if (self.group is not None or not isinstance(self.plan, CudaPlan) or out_of_env("PINN_OWN_NCCL")):
            return None
        lib, dev = self.plan.lib, self.flat.device
        uid = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (C.c_ubyte * 128)()
            _capi.check(lib.pinn_nccl_unique_id(buf), "pinn_nccl_unique_id")
            uid = torch.tensor(list(buf), dtype=torch.uint8)
        uid = uid.to(dev)
        self._dist.broadcast(uid, src=0)
        raw = (C.c_ubyte * 128)(*uid.cpu().tolist())
        comm = C.c_void_p()
        _capi.check(lib.pinn_comm_create(raw, self.world, self.rank, dev.index if dev.index is not None else torch.cuda.current_device(),
                                         C.byref(comm)), "pinn_comm_create")
        self._comm = comm
        return comm

    def loss_and_grad_device(self):
        """One training-step evaluation, everything left on the device:
        returns (grad [P] view, sumsq [T] in table order)."""
        out = self._reduce(self.plan.loss_and_grad(self.flat))
        return out[: self.compiled.n_params], self.plan.to_table_order(out)

    def training_step(self, optimizer):
        """One full training step, asynchronous: loss step (+ all-reduce) + optimiser update.
        Returns the device vector of per-term sums of squares (table order).

        The step is the same launch sequence every time (fixed point sets, device-side Adam step number), so after three
        eager steps it is captured once in a CUDA graph and replayed: the small configs (Poiseuille 10 k, Colliding 100 k
        points) are launch-bound otherwise.  With several ranks the NCCL all-reduce of the [P+T] vector is captured with it
        (every rank captures at the same step), which removes the eager launch gaps around the 9 KB collective.
        ``PINN_CUDA_GRAPH=0`` disables the replay, ``PINN_CUDA_GRAPH_DIST=0`` only the multi-rank one."""
        if self._graph is not None and optimizer is self._graph_opt and not self.plan.timing_enabled:
            self._graph.replay()
            optimizer.t += 1
            return self._graph_sumsq
        if self._graph_eligible(optimizer):
            self._eager_steps += 1
            if self._eager_steps > 3:
                return self._capture_step(optimizer)
        grad, sumsq = self.loss_and_grad_device()
        optimizer.apply(self.flat, grad)
        return sumsq

    def _graph_eligible(self, optimizer) -> bool:
        return (self.flat.is_cuda and (self.world == 1 or os.environ.get("PINN_CUDA_GRAPH_DIST", "1") != "0")
                and isinstance(optimizer, Adam) and isinstance(self.plan, CudaPlan)
                and not self.plan.timing_enabled and os.environ.get("PINN_CUDA_GRAPH", "1") != "0"
                and (self._graph_opt is None or self._graph_opt is optimizer))

    def _capture_step(self, optimizer):
        torch.cuda.synchronize(self.flat.device)
        g = torch.cuda.CUDAGraph()
        # several ranks: torch's NCCL watchdog thread may touch the CUDA API while this thread captures
        with torch.cuda.graph(g, capture_error_mode="thread_local" if self.world > 1 else "global"):
            grad, sumsq = self.loss_and_grad_device()
            optimizer.apply(self.flat, grad)
        optimizer.t -= 1                 # capture launches nothing: the step is taken by the replay below
        self._graph, self._graph_opt, self._graph_sumsq = g, optimizer, sumsq
        g.replay()
        optimizer.t += 1
        return sumsq

    def evaluate(self):
        """(total loss, per-train-term values, flat gradient [P] on device)."""
        grad, sumsq = self.loss_and_grad_device()
        total, train_vals, _ = assemble_losses(self.compiled, sumsq.detach().double().cpu().numpy())
        return total, train_vals, grad

    def evaluate_host(self, theta: np.ndarray):
        """(total, float64 gradient) at host parameters ``theta`` -- the objective SciPy-style drivers call thousands of
        times.  One pinned H2D copy of the parameters, one loss step, ONE pinned D2H copy of the [P+T] result
        (``evaluate`` issues several small kernels and two synchronising reads)."""
        P = self.compiled.n_params
        if not (self.flat.is_cuda and isinstance(self.plan, CudaPlan)):
            self.flat.copy_(torch.as_tensor(theta, dtype=torch.float32))
            total, _, grad = self.evaluate()
            return float(total), grad.detach().double().cpu().numpy()
        if self._h_theta is None:
            self._h_theta = torch.empty(P, dtype=torch.float32, pin_memory=True)
            self._h_out = torch.empty(self.plan.out.numel(), dtype=torch.float32, pin_memory=True)
            self._perm_np = np.asarray(self.plan._perm, dtype=np.int64)
        self._h_theta.numpy()[:] = theta
        self.flat.copy_(self._h_theta, non_blocking=True)
        out = self._reduce(self.plan.loss_and_grad(self.flat))
        self._h_out.copy_(out, non_blocking=True)
        torch.cuda.current_stream(self.flat.device).synchronize()
        o = self._h_out.numpy()
        sums = np.zeros(max(1, self.compiled.n_out_terms))
        if self.plan.n_kernel_terms:
            sums[self._perm_np] = o[P:P + self.plan.n_kernel_terms]
        total, _, _ = assemble_losses(self.compiled, sums)
        return float(total), o[:P].astype(np.float64)

    def evaluate_all(self):
        """Forward-only values of train and test terms (log points)."""
        out = self._reduce(self.plan.loss_only(self.flat))
        sumsq = self.plan.to_table_order(out).detach().double().cpu().numpy()
        return assemble_losses(self.compiled, sumsq)

    # ---- history ------------------------------------------------------------------------------
    def begin_round(self, name: str) -> None:
        h = self.history
        if h["log"]["iter"]:
            self.iteration = h["log"]["iter"][-1] + 1     # iteration_start == [0, 101] in the saved runs
        h["log_rounds"]["rounds"].append(name)
        h["log_rounds"]["iteration_start"].append(self.iteration)
        self._iter_round = 0

    def log_state(self) -> float:
        total, train_vals, test_vals = self.evaluate_all()
        h = self.history
        h["log"]["iter"].append(self.iteration)
        h["log"]["round"].append(len(h["log_rounds"]["rounds"]))
        h["log"]["iter_round"].append(self._iter_round)
        h["log"]["loss_global"].append(total)
        for l, v in zip(self.losses, train_vals):
            h["losses"][l.name]["log"].append(v)
        for l, v in zip(self.losses_test, test_vals):
            h["losses_test"][l.name]["log"].append(v)
        for cb in self.callbacks:
            cb(self, self._iter_round)
        return total

    def step_done(self) -> None:
        self.iteration += 1
        self._iter_round += 1
        if self._iter_round % LOG_STRIDE == 0:
            self.log_state()

    def save_history(self, path: str) -> None:
        if self.rank == 0:
            with open(path, "w") as fh:
                json.dump(self.history, fh, indent=2)


# ------------------------------------------------------------------------------------------------
# optimisers and minimize
# ------------------------------------------------------------------------------------------------

class Adam:
    """tf.keras.optimizers.Adam(learning_rate=1e-2): beta_1 .9, beta_2 .999, epsilon 1e-7 (Keras 2.7
    defaults), update  theta -= lr*sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps)."""

    name = "Adam"

    def __init__(self, learning_rate: float = 1e-3, beta_1: float = 0.9, beta_2: float = 0.999, epsilon: float = 1e-7):
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta_1, beta_2, epsilon
        self.m = self.v = None
        self.t = 0

    def apply(self, flat: torch.Tensor, grad: torch.Tensor) -> None:
        import ctypes as C
        if self.m is None:
            self.m, self.v = torch.zeros_like(flat), torch.zeros_like(flat)
            if flat.is_cuda:      # step number on the device: the update launches with the same arguments every step
                self.step_dev = torch.full((1,), self.t, dtype=torch.int64, device=flat.device)
        self.t += 1
        if flat.is_cuda:
            lib = _capi.load()
            stream = C.c_void_p(torch.cuda.current_stream(flat.device).cuda_stream)
            _capi.check(lib.pinn_adam_step_dev(C.c_void_p(flat.data_ptr()), C.c_void_p(grad.data_ptr()),
                                               C.c_void_p(self.m.data_ptr()), C.c_void_p(self.v.data_ptr()),
                                               flat.numel(), self.lr, self.b1, self.b2, self.eps,
                                               C.c_void_p(self.step_dev.data_ptr()), stream),
                        "pinn_adam_step_dev")
        else:  # host tensors: only reached with an injected (test) engine
            self.m.mul_(self.b1).add_(grad, alpha=1 - self.b1)
            self.v.mul_(self.b2).addcmul_(grad, grad, value=1 - self.b2)
            step = self.lr * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
            flat.addcdiv_(self.m, self.v.sqrt().add_(self.eps), value=-step)


class optimizers:
    Adam = Adam


def minimize(pb: OptimizationProblem, backend: str, optimizer, num_epochs: int) -> None:
    if backend == "keras":
        pb.begin_round(f"keras_{getattr(optimizer, 'name', type(optimizer).__name__)}")
        pb.log_state()
        for _ in range(num_epochs):
            pb.training_step(optimizer)
            pb.step_done()
    elif backend == "scipy":
        import scipy.optimize
        method = str(optimizer)
        pb.begin_round(f"scipy_{method}")
        pb.log_state()
        # Every saved BFGS round of the reference ends one log short of `epochs` (History_Loss.json of Cavity_Steady,
        # Colliding_Flow, Poiseuille_Flow: iter_round 9990 for epochs = 10000; Coronary_Flow 29990 for 30000), the saved
        # L-BFGS-B rounds (Examples_Old) at iter_round == epochs: [inferred] nisaba runs BFGS for epochs - 1 iterations.
        maxiter = int(num_epochs) - 1 if method.upper() == "BFGS" else int(num_epochs)
        mode = os.environ.get("PINN_BFGS", "device")
        if method.upper() == "BFGS" and mode == "device" and pb.flat.is_cuda and isinstance(pb.plan, CudaPlan):
            # SciPy's BFGS algorithm and line search with the quasi-Newton algebra in hand-written kernels on the device
            # (scipy.optimize.minimize spends 133 ms per iteration in two P^3 products for P = 2307): bfgs.py, csrc/bfgs.cuh
            from .bfgs import minimize_bfgs_device
            pb.last_result = minimize_bfgs_device(pb, maxiter=maxiter, callback=pb.step_done)
            return

        def fun(theta: np.ndarray):
            return pb.evaluate_host(theta)

        def cb(theta):
            # the history logs the ACCEPTED iterate (pb.flat may hold the last line-search trial point)
            pb.flat.copy_(torch.as_tensor(np.asarray(theta), dtype=torch.float32))
            pb.step_done()

        x0 = pb.flat.detach().double().cpu().numpy()
        if method.upper() == "BFGS" and mode != "scipy":
            from .bfgs import minimize_bfgs      # same algorithm on host vectors (any engine)
            res = minimize_bfgs(fun, x0, maxiter=maxiter, callback=cb, device=pb.flat.device)
        else:
            res = scipy.optimize.minimize(fun, x0, jac=True, method=method, callback=cb, options={"maxiter": maxiter})
        pb.last_result = res
        pb.flat.copy_(torch.as_tensor(res.x, dtype=torch.float32))
    else:
        raise ValueError(f"unknown backend {backend!r} (expected 'keras' or 'scipy')")


# ------------------------------------------------------------------------------------------------
# utils
# ------------------------------------------------------------------------------------------------

class utils:
    @staticmethod
    def load_json(path: str):
        with open(path) as fh:
            return json.load(fh)

    @staticmethod
    def plot_history(path: str) -> None:
        """matplotlib is not available in this image; the JSON is what the reference's own
        ``plot_loss`` consumes (cavity_steady.py:339-362)."""
        return None

    class HistoryPlotCallback:
        """Every ``frequency`` iterations of a round: dump the history JSON (and, in the reference,
        a PNG) -- cavity_steady.py:243-245."""

        def __init__(self, frequency=100, gui=False, filename=None, filename_history=None):
            self.frequency, self.filename, self.filename_history = frequency, filename, filename_history

        def __call__(self, pb: OptimizationProblem, iter_round: int) -> None:
            if self.filename_history and iter_round % self.frequency == 0:
                pb.save_history(self.filename_history)
