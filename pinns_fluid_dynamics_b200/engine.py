"""Compilation of a loss table into point-set descriptors and the CUDA plan that evaluates them.

``compile_problem`` is pure host logic (numpy): it groups the loss terms by point set, shards every
point set over the ranks and computes the scale factors -- the part of nisaba's
``OptimizationProblem`` that is not arithmetic.  ``CudaPlan`` binds the result to libpinnstep.so; it
is the only evaluator the product ships (no CPU fallback).

Sharding: every point set is cut into ``world`` contiguous chunks, the first
``n mod world`` ranks get one extra row; a set smaller than ``world`` leaves the last ranks empty.
Kernels return raw sum r^2 per term and the gradient of sum_t w_t/(nu_t N_t) sum r^2 with N_t the
GLOBAL count, so a plain SUM all-reduce of the [P+T] vector is exact.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _capi
from .residuals import MAX_CH, MAX_OUT, PointSet, ResidualForm

MAX_TERMS_PER_SET = _capi.MAX_TERMS_PER_SET


@dataclass
class CompiledTerm:
    name: str
    form: ResidualForm
    weight: float
    normalization: float
    train: bool
    n_global: int
    out_index: int = -1            # position in the [T] tail of the output vector (-1: identically zero)

    @property
    def abs_mean(self) -> bool:
        """|mean r| reduction (ns.Loss): the output slot carries sum r instead of sum r^2"""
        return self.form.reduction == "abs_mean"


@dataclass
class CompiledSet:
    pointset: PointSet
    start: int                     # first global row owned by this rank
    stop: int
    deriv_order: int
    terms: List[CompiledTerm] = field(default_factory=list)

    @property
    def n_local(self) -> int:
        return self.stop - self.start


@dataclass
class CompiledProblem:
    mlp: Tuple[int, int, int, int]           # (d, H, L, O)
    sets: List[CompiledSet]
    terms: List[CompiledTerm]                # train terms first (table order), then test terms
    n_params: int
    rank: int
    world: int

    @property
    def n_out_terms(self) -> int:
        return sum(1 for t in self.terms if t.out_index >= 0)


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def mlp_shape_from_variables(shapes: Sequence[Tuple[int, ...]]) -> Tuple[int, int, int, int]:
    """(d, H, L, O) from Keras-ordered variable shapes [K1, b1, ..., K_{L+1}, b_{L+1}]."""
    if len(shapes) < 4 or len(shapes) % 2:
        raise ValueError("variables must be [K1, b1, ..., K_out, b_out] with at least one hidden layer")
    kernels, biases = shapes[0::2], shapes[1::2]
    d, H = kernels[0]
    L = len(kernels) - 1
    O = kernels[-1][1]
    for i, (k, b) in enumerate(zip(kernels, biases)):
        exp_in = d if i == 0 else H
        exp_out = O if i == L else H
        if tuple(k) != (exp_in, exp_out) or tuple(b) != (exp_out,):
            raise ValueError(f"layer {i}: expected kernel {(exp_in, exp_out)} / bias {(exp_out,)}, got {k} / {b}")
    return int(d), int(H), int(L), int(O)


def param_count(d: int, H: int, L: int, O: int) -> int:
    return d * H + H + (L - 1) * (H * H + H) + H * O + O


def compile_problem(var_shapes, losses, losses_test=(), rank: int = 0, world: int = 1) -> CompiledProblem:
    """``losses`` / ``losses_test``: objects with .name, .form (ResidualForm), .weight, .normalization."""
    d, H, L, O = mlp_shape_from_variables(var_shapes)
    if O > MAX_OUT:
        raise ValueError(f"at most {MAX_OUT} network outputs are supported")
    terms: List[CompiledTerm] = []
    for train, table in ((True, losses), (False, losses_test)):
        for l in table:
            f = l.form
            if f.pointset.dim != d:
                raise ValueError(f"term {l.name}: point set has dim {f.pointset.dim}, network input is {d}")
            for (o, c), v in f.coef.items():
                if v != 0.0 and (o >= O or c >= 1 + d + 2):
                    raise ValueError(f"term {l.name}: coefficient on output {o} / channel {c} out of range")
            if f.conv != 0.0 and O < 2:
                raise ValueError(f"term {l.name}: convective product needs outputs 0 and 1")
            terms.append(CompiledTerm(l.name, f, float(l.weight), float(l.normalization), train, f.pointset.n))
    # group by point set, first-appearance order; split groups larger than the ABI allows
    by_set: Dict[int, List[CompiledTerm]] = {}
    order: List[PointSet] = []
    for t in terms:
        if t.form.identically_zero:
            continue
        ps = t.form.pointset
        if ps.uid not in by_set:
            by_set[ps.uid] = []
            order.append(ps)
        by_set[ps.uid].append(t)
    sets: List[CompiledSet] = []
    nxt = 0
    for ps in order:
        group = by_set[ps.uid]
        start, stop = shard_bounds(ps.n, rank, world)
        if any(t.abs_mean for t in group):
            # the sign of a |mean| term comes from a forward pre-pass over its whole point set (include/pinnstep.h):
            # such a set is not sharded, rank 0 owns it
            start, stop = (0, ps.n) if rank == 0 else (0, 0)
        for i in range(0, len(group), MAX_TERMS_PER_SET):
            chunk = group[i:i + MAX_TERMS_PER_SET]
            cs = CompiledSet(ps, start, stop, max(t.form.deriv_order() for t in chunk), chunk)
            sets.append(cs)
    for t in terms:  # output order = table order (train first), independent of grouping
        if not t.form.identically_zero:
            t.out_index = nxt
            nxt += 1
    return CompiledProblem((d, H, L, O), sets, terms, param_count(d, H, L, O), rank, world)


def assemble_losses(cp: CompiledProblem, sumsq: np.ndarray) -> Tuple[float, List[float], List[float]]:
    """(total, train values, test values) from the GLOBAL sums of squares.
    value_t = sum r^2 / (N_t * normalization_t) (|sum r| / (N_t * normalization_t) for an ns.Loss |mean| term);
    total = sum_train weight_t * value_t
    (nisaba semantics evidenced by History_Loss.json).  An empty point set gives
    NaN, like the reference's mean over an empty tensor (quirk Q3)."""
    train_vals, test_vals = [], []
    total = 0.0
    for t in cp.terms:
        if t.out_index < 0:
            v = 0.0
        elif t.n_global == 0:
            v = float("nan")
        elif t.abs_mean:      # ns.Loss over |mean(roots)|: the slot carries sum r
            v = abs(float(sumsq[t.out_index])) / (t.n_global * t.normalization)
        else:
            v = float(sumsq[t.out_index]) / (t.n_global * t.normalization)
        if t.train:
            train_vals.append(v)
            total += t.weight * v
        else:
            test_vals.append(v)
    return total, train_vals, test_vals


class CudaPlan:
    """Device-side evaluator: owns the device copies of the local point/rhs shards and the
    ``pinn_plan``.  Raises if CUDA or the library is unavailable."""

    def __init__(self, cp: CompiledProblem, device=None):
        import torch
        if not torch.cuda.is_available():
            raise _capi.PinnLibraryError("CUDA device required: the PINN loss step has no CPU fallback")
        self.lib = _capi.load()
        self.cp = cp
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise _capi.PinnLibraryError(f"CUDA device required, got {self.device}: the PINN loss step has no CPU fallback")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        # every local point / rhs shard lives in ONE device arena (views below), mirrored by one host arena:
        # the end-to-end path moves all inputs of a step with a single host->device copy
        hosts, where = [], {}
        for si, cs in enumerate(cp.sets):
            key = ("pts", cs.pointset.uid)
            if cs.n_local and key not in where:
                where[key] = len(hosts)
                hosts.append(np.ascontiguousarray(cs.pointset.points[cs.start:cs.stop]))
            for ti, t in enumerate(cs.terms):
                rhs = t.form.rhs_array()
                if rhs is not None and cs.n_local:
                    where[("rhs", si, ti)] = len(hosts)
                    hosts.append(np.ascontiguousarray(rhs[cs.start:cs.stop]))
        offs, total = [], 0
        for h in hosts:
            offs.append(total)
            total += (h.nbytes + 255) // 256 * 256
        self._host_arena = np.zeros(max(total, 256), dtype=np.uint8)
        for h, o in zip(hosts, offs):
            self._host_arena[o:o + h.nbytes] = h.view(np.uint8).reshape(-1)
        self._arena = torch.from_numpy(self._host_arena).to(self.device)
        self._input_bytes = sum(h.nbytes for h in hosts)
        views = [self._arena[o:o + h.nbytes].view(torch.from_numpy(h).dtype).view(h.shape) for h, o in zip(hosts, offs)]
        self._pinned = None
        descs = (_capi.PointSetDesc * max(1, len(cp.sets)))()
        for si, cs in enumerate(cp.sets):
            ds = descs[si]
            ds.points_dev = views[where[("pts", cs.pointset.uid)]].data_ptr() if cs.n_local else None
            ds.n_local = cs.n_local
            ds.n_terms = len(cs.terms)
            ds.deriv_order = cs.deriv_order
            for ti, t in enumerate(cs.terms):
                td = ds.terms[ti]
                m = t.form.coef_matrix()
                for o in range(MAX_OUT):
                    for c in range(MAX_CH):
                        td.coef[o][c] = float(m[o, c])
                td.conv, td.conv_k, td.rhs_scale = float(t.form.conv), int(t.form.conv_k), float(t.form.rhs_scale)
                if ("rhs", si, ti) in where:
                    td.rhs_dev = views[where[("rhs", si, ti)]].data_ptr()
                else:
                    td.rhs_dev = None
                td.weight, td.normalization = t.weight, t.normalization
                td.n_global, td.train = t.n_global, 1 if t.train else 0
                td.kind = 1 if t.abs_mean else 0
        d, H, L, O = cp.mlp
        mlp = _capi.MlpDesc(d, H, L, O)
        handle = C.c_void_p()
        _capi.check(self.lib.pinn_plan_create(C.byref(mlp), descs, len(cp.sets), self.device.index or 0,
                                              C.byref(handle)), "pinn_plan_create")
        self.handle = handle
        # kernel term order is set order / term order; map to table order
        self._perm = [t.out_index for cs in cp.sets for t in cs.terms]
        self.n_kernel_terms = len(self._perm)
        self.P = cp.n_params
        self.out = torch.zeros(self.P + max(1, self.n_kernel_terms), dtype=torch.float32, device=self.device)
        self._perm_t = torch.tensor(self._perm, dtype=torch.long, device=self.device)
        self.engine = self.lib.pinn_plan_engine(handle).decode()

    def _stream(self):
        import torch
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def loss_and_grad(self, params_flat):
        """Enqueue one step; returns the device vector [P + T] (gradient, then per-term sum r^2 in
        KERNEL term order -- use ``to_table_order``)."""
        self._check_params(params_flat)
        _capi.check(self.lib.pinn_loss_and_grad(self.handle, C.c_void_p(params_flat.data_ptr()),
                                                C.c_void_p(self.out.data_ptr()), self._stream()),
                    "pinn_loss_and_grad")
        return self.out

    def _check_params(self, params_flat):
        import torch
        if not (params_flat.is_cuda and params_flat.dtype == torch.float32 and params_flat.is_contiguous()):
            raise ValueError("the parameter vector must be a contiguous FP32 CUDA tensor")
        if params_flat.device != self.device:
            raise ValueError(f"parameters live on {params_flat.device}, the plan (point sets, workspace) on {self.device}")

    def loss_only(self, params_flat):
        self._check_params(params_flat)
        _capi.check(self.lib.pinn_loss(self.handle, C.c_void_p(params_flat.data_ptr()),
                                       C.c_void_p(self.out.data_ptr()), self._stream()), "pinn_loss")
        return self.out

    def to_table_order(self, out):
        """[T_table] tensor of sum r^2 indexed by CompiledTerm.out_index."""
        import torch
        res = torch.zeros(max(1, self.cp.n_out_terms), dtype=out.dtype, device=out.device)
        if self.n_kernel_terms:
            res[self._perm_t] = out[self.P:self.P + self.n_kernel_terms]
        return res

    def last_launch_count(self) -> int:
        return int(self.lib.pinn_plan_last_launch_count(self.handle))

    # ---- host-resident inputs (end-to-end path: data arrives in host memory every step) -------
    def pin_host_inputs(self) -> int:
        """Stage every local point / rhs shard in pinned host memory; returns the byte count one
        ``upload_inputs`` moves."""
        import torch
        if self._pinned is None:
            # allocate page-locked memory directly: measured 54 GB/s H2D, against 13-22 GB/s from Tensor.pin_memory()
            self._pinned = torch.empty(self._host_arena.shape[0], dtype=torch.uint8, pin_memory=True)
            self._pinned.copy_(torch.from_numpy(self._host_arena))
        return int(self._pinned.numel())

    def upload_inputs(self) -> None:
        """Host -> device copy of all inputs into the buffers the plan points at: one asynchronous copy of the
        pinned host arena on the current stream (serial with the step that follows)."""
        if self._pinned is None:
            self.pin_host_inputs()
        self._arena.copy_(self._pinned, non_blocking=True)

    # double-buffered variant: the inputs of step i+1 travel while step i computes
    def prefetch_inputs(self) -> None:
        """Start the host -> device copy of the NEXT step's inputs on a side stream into a staging buffer.  Two pinned
        host arenas alternate, so the host may refill one while the copy engine reads the other.  (With memory from
        ``Tensor.pin_memory()`` this overlap was pathological -- 12 ms per copy under a kernel that fills every SM;
        with directly page-locked allocations it costs nothing: tools/h2d_overlap_probe.py, 1.61 -> 1.49 ms per step.)"""
        import torch
        if self._pinned is None:
            self.pin_host_inputs()
        if self._stage is None:
            self._stage = torch.empty_like(self._arena)
            self._pinned2 = torch.empty(self._pinned.numel(), dtype=torch.uint8, pin_memory=True)
            self._pinned2.copy_(self._pinned)
            self._side = torch.cuda.Stream(device=self.device)
            self._ready, self._consumed = torch.cuda.Event(), torch.cuda.Event()
            self._consumed.record(torch.cuda.current_stream(self.device))
            self._flip = 0
        src = self._pinned if self._flip == 0 else self._pinned2
        self._flip ^= 1
        with torch.cuda.stream(self._side):
            self._side.wait_event(self._consumed)        # the staging buffer has been consumed by commit_inputs
            self._stage.copy_(src, non_blocking=True)
            self._ready.record(self._side)

    def commit_inputs(self) -> None:
        """Make the prefetched inputs current: wait for the side-stream copy and move the staging buffer into the
        plan's arena (device-to-device, on the current stream, before the step that uses it)."""
        import torch
        main = torch.cuda.current_stream(self.device)
        main.wait_event(self._ready)
        self._arena.copy_(self._stage, non_blocking=True)
        self._consumed.record(main)

    _stage = None

    timing_enabled = False

    def enable_timing(self, on: bool = True) -> None:
        _capi.check(self.lib.pinn_plan_enable_timing(self.handle, 1 if on else 0), "pinn_plan_enable_timing")
        self.timing_enabled = bool(on)

    def kernel_time_ms(self, deriv_order: int = 2) -> float:
        ms = C.c_float()
        _capi.check(self.lib.pinn_plan_kernel_time_ms(self.handle, deriv_order, C.byref(ms)), "pinn_plan_kernel_time_ms")
        return float(ms.value)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.pinn_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
