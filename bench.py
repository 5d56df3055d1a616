#!/usr/bin/env python
"""bench.py -- collocation points per second per PINN training step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config Cavity_Steady] [--impl ours|reference]

One "step" = forward jets + Navier-Stokes residuals + weighted mean-square loss + parameter
gradients (+ all-reduce of the [P+T] vector when N>1) + Adam update, on the BASELINE workload
(default: Cavity_Steady, 1 000 000 synthetic collocation points PER GPU -> weak scaling, plus the
script's 4x1000 boundary, 100 velocity and 1 pressure fitting points).

Printed keys (one JSON line on rank 0):
  value      whole-job pts/s with inputs resident in HBM (device-timed per step, L2 flushed between steps)
  e2e        the same through the public facade with HOST inputs: every step copies all point / target
             arrays from pinned host memory, runs the step, and reads the per-term sums back (the host reads step i
             after it has enqueued step i+1, so the device never waits for the host)
  roofline   the collocation kernel against the tensor roofline of its precision mode (3xTF32: the MEASURED tcgen05 kind::tf32
             rate / 3, profiles/tf32_peak_r01.json -- MEASURED_PEAKS.json holds no TF32 figure), and, in `two_pipe`, against the
             floors of BOTH pipes it needs: the tensor pipe (MMAs issued x measured cycles per MMA) and the FMA / issue pipes
             (FP32-pipe and total warp instructions per point of the shipped kernel, ncu capture named in `source`)
  cpu_baseline   oracle/reference_step.py (torch float64 nested autodiff, the reference's step structure) timed on this
             box's host cores on the SAME workload when ~25 s of CPU time allow it, else on the largest sample that fits
  other_configs  device-timed step of the four other BASELINE.json configs (N = 1: their own sizes; N > 1: sharded)
  strong_scaling (N > 1) the 1 000 000-point Cavity_Steady workload sharded over the N GPUs
`--impl reference` times that CPU restatement alone (the real nisaba/TensorFlow stack cannot be
installed: see DESIGN.md) on the same config/metric/unit.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CPU_SAMPLE_PDE = 20_000          # timing probe of the CPU arm; the sample grows from here up to the full workload


def flops_per_point(d, H, L, O, C):
    """matmul FLOPs only, 2 per MAC, step = 3 x forward."""
    return 3 * (2 * d * H + (L - 1) * C * 2 * H * H + C * 2 * H * O)


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons during the timed regions (NVML, 10 ms period)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self.active = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            if self.active.is_set():
                try:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    try:
                        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for bit, name in names.items():
                        if r & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
            time.sleep(0.01)

    def stop(self):
        self._stop.set()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def _cpu_step_fn(config: str, n_pde: int):
    """(one_step(t), variables) of oracle/reference_step.py on `n_pde` collocation points + every boundary / fit set."""
    import torch
    from oracle import reference_step
    from pinns_fluid_dynamics_b200 import problems

    data = problems.build_baseline_config(config, seed=1, PDE=n_pde)
    var = reference_step.glorot_uniform_variables(data.dim, data.hidden, data.out_dim, seed=3)
    pb = reference_step.build(data, var)
    m = [torch.zeros_like(v) for v in pb.variables]
    v2 = [torch.zeros_like(v) for v in pb.variables]

    def one_step(t):
        _, _, grad = pb.loss_and_grad()
        off = 0
        with torch.no_grad():   # Adam(1e-2) like cavity_steady.py:246
            for i, p in enumerate(pb.variables):
                g = grad[off:off + p.numel()].view_as(p); off += p.numel()
                m[i].mul_(0.9).add_(g, alpha=0.1)
                v2[i].mul_(0.999).addcmul_(g, g, value=0.001)
                step = 1e-2 * (1 - 0.999 ** t) ** 0.5 / (1 - 0.9 ** t)
                p.addcdiv_(m[i], v2[i].sqrt().add_(1e-7), value=-step)
    return one_step


def cpu_reference_step_rate(config: str, steps: int, warmup: int, budget_s: float, full_pde: int):
    """Time oracle/reference_step.py (the reference's step structure, torch float64, all host threads).  The sample is the
    FULL workload (`full_pde` collocation points) when (steps + warmup) steps of it fit `budget_s` seconds of CPU time and
    host memory, else the largest power-of-two fraction that does; a 20 000-point probe sets the scale.
    Returns (pts/s, ms/step, cores, sample text, sample points)."""
    import torch

    torch.set_num_threads(os.cpu_count() or 1)
    probe_n = min(CPU_SAMPLE_PDE, full_pde)
    step = _cpu_step_fn(config, probe_n)
    step(1)
    t0 = time.perf_counter()
    step(2)
    t_probe = time.perf_counter() - t0
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 32 << 30
    n = full_pde
    # 40 KB of tape per point; the step is up to ~3x slower per point once the tape leaves the caches (200 k points), so the
    # probe only rules out sizes that are hopeless (3x over budget even at the probe's rate) -- the decision is then made on
    # a MEASURED step at the candidate size, which also serves as its first warm-up step
    while n > probe_n and ((steps + warmup) * t_probe * (n / probe_n) > 3.0 * budget_s or n * 40_000 > 0.5 * avail):
        n //= 2
    n = max(n, probe_n)
    done_warm = 0
    while True:
        if n != probe_n:
            step = _cpu_step_fn(config, n)
        t0 = time.perf_counter()
        step(1)
        t_one = time.perf_counter() - t0
        done_warm = 1
        if n <= probe_n or (steps + warmup) * t_one <= budget_s:
            break
        n = max(n // 2, probe_n)
        if n == probe_n:
            step = _cpu_step_fn(config, n)
            done_warm = 0
            break
    for i in range(done_warm, warmup):
        step(i + 1)
    t0 = time.perf_counter()
    for i in range(steps):
        step(warmup + i + 1)
    dt = time.perf_counter() - t0
    what = "the full workload" if n == full_pde else f"{n} of the {full_pde} collocation points"
    sample = (f"{config}: {what} + all boundary/fit sets, {steps} steps after {warmup} warm-up, "
              f"torch {torch.__version__} float64, {torch.get_num_threads()} threads")
    return n * steps / dt, dt / steps * 1e3, torch.get_num_threads(), sample, n


def workload_text(config, per_gpu, n_global, data, d, H, L, O, n_params, n_terms):
    return (f"{config}: {per_gpu} uniform collocation points per GPU ({n_global} total), "
            f"{data.options.n_pts['BC']}x4 boundary, {data.options.n_pts['Vel']} velocity + "
            f"{data.options.n_pts['Pres']} pressure fitting points, tanh MLP "
            f"{d}-{H}x{L}-{O} ({n_params} parameters), {n_terms} loss terms")


def run_reference(args):
    """The reference arm: the CPU restatement of the nisaba / TensorFlow step on this box's host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from pinns_fluid_dynamics_b200 import loss_tables, problems
    full = args.pde_per_gpu or problems.BASELINE_CONFIGS[args.config].get("PDE", 200)
    warm = max(args.warmup, 1)
    value, ms, cores, sample, n = cpu_reference_step_rate(args.config, args.steps, warm, budget_s=240.0, full_pde=full)
    data = problems.build_baseline_config(args.config, seed=1, PDE=1000)
    losses, _ = loss_tables.build_loss_table(data, faithful=data.name.startswith("cavity"))
    L = len(data.hidden)
    n_params = data.dim * data.hidden[0] + data.hidden[0] + (L - 1) * (data.hidden[0] ** 2 + data.hidden[0]) + data.hidden[0] * data.out_dim + data.out_dim
    line = {
        "impl": "reference", "metric": "collocation_points_per_second_per_training_step", "value": value,
        "unit": "pts/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": warm, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_text(args.config, full, full, data, data.dim, data.hidden[0], L, data.out_dim, n_params, len(losses)),
                   "cpu_sample_points_per_step": n, "full_workload": n == full},
        "cpu_baseline": {"value": value, "unit": "pts/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "pts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU restatement of the reference's nisaba/TensorFlow step (oracle/reference_step.py); "
                "TensorFlow 2.7 and nisaba are not installable in this image",
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import pinns_fluid_dynamics_b200 as ns
    from pinns_fluid_dynamics_b200 import loss_tables, problems

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    per_gpu = args.pde_per_gpu or problems.BASELINE_CONFIGS[args.config].get("PDE", 200)   # Poisson: 200 (poisson_misto.py:49)
    data = problems.build_baseline_config(args.config, seed=1, PDE=per_gpu * world)
    model = ns.TanhMLP(data.dim, data.hidden, data.out_dim, device=dev, seed=3)
    # benchmark the full Navier-Stokes residual: in-tape mass term everywhere
    faithful = data.name.startswith("cavity")
    losses, ltest = loss_tables.build_loss_table(data, faithful=faithful)
    pb = ns.OptimizationProblem(model.variables, losses, ltest)
    opt = ns.Adam(learning_rate=1e-2)
    plan = pb.plan
    d, H, L, O = pb.compiled.mlp
    n_local_pde = sum(cs.n_local for cs in pb.compiled.sets if cs.deriv_order == 2 and cs.pointset.name == "PDE")
    n_global_pde = per_gpu * world

    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2
    sampler = ClockSampler(local_rank)
    sampler.start()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ---------------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        pb.training_step(opt)
    # on one GPU the facade captures the step in a CUDA graph after three eager steps: that one-off capture (a host-side
    # pause of several ms) belongs to the warm-up, whatever --warmup says
    extra = 0
    while getattr(pb, "_graph", None) is None and pb._graph_eligible(opt) and extra < 8:
        pb.training_step(opt)
        extra += 1
    launches_per_step = plan.last_launch_count() + 1   # + Adam kernel (it also advances the device step counter)
    if world > 1 and getattr(pb, "_p2p", None) is not None:
        launches_per_step += 1                         # + the one-kernel peer-memory all-reduce
    barrier()

    # ---- value: inputs resident in HBM, device-timed per step, L2 flushed between steps -----------
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    sampler.active.set()
    for e0, e1 in ev:
        flush.fill_(1.0)
        e0.record()
        pb.training_step(opt)
        e1.record()
    barrier()
    sampler.active.clear()
    ms_dev = sum(e0.elapsed_time(e1) for e0, e1 in ev)
    t = torch.tensor([ms_dev], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev = float(t.item())
    value = n_global_pde * args.steps / (ms_dev * 1e-3)

    # ---- roofline: the collocation kernel alone, CUDA events around the launch inside the C ABI ----
    plan.enable_timing(True)
    kms = []
    for _ in range(min(args.steps, 50)):
        flush.fill_(1.0)
        pb.training_step(opt)
        kms.append(plan.kernel_time_ms(2))
    plan.enable_timing(False)
    k_ms = float(np.mean(kms))
    C = 3 + d
    flop_launch = n_local_pde * flops_per_point(d, H, L, O, C)
    achieved = flop_launch / (k_ms * 1e-3) * 1e-12
    traffic = None
    extra = {}
    if plan.engine == "layered_tf32x3":
        # hidden-layer contractions on tcgen05 kind::tf32, 3 MMA passes per algorithmic product (hi/lo split)
        bound, peak, peak_src = "tensor", 368.4, "fallback"
        kernel_name = ("collocation pipeline of the layered engine: tc_layer1, tc_layer<fwd> x(L-1), tc_out_layer, "
                       "(tc_wgrad, tc_layer<bwd>) x(L-1), tc_layer1_grad per batch; timed as one event pair")
        try:
            with open(os.path.join(ROOT, "profiles", "tf32_peak_r01.json")) as fh:
                pk = json.load(fh)
            peak = pk["tf32_mma_tflops_n256"] / pk["passes_per_product"]
            peak_src = ("profiles/tf32_peak_r01.json: measured tcgen05 kind::tf32 MMA rate (tools/tc_probe.cu) / 3 passes; "
                        "MEASURED_PEAKS.json holds no TF32 figure")
        except Exception:
            pass
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic_tc_r01.json")) as fh:
                traffic = json.load(fh)["dram_bytes_per_point"] * n_local_pde
        except Exception:
            pass
    elif plan.engine == "fused_tcgen05":
        # hidden-layer forward GEMMs on tcgen05 kind::tf32 (3 passes per product, operands in tensor memory), adjoint GEMMs and
        # weight gradient as bf16-pair kind::f16 MMAs; `peak` is the chip's 3xTF32 tensor roofline.  The kernel also needs the FMA /
        # ALU pipes for the tanh-jet math and the operand splits: `two_pipe` holds the floors of both.
        bound, peak, peak_src = "tensor", 368.4, "fallback"
        kernel_name = ("fused_tc_kernel<D=2,O=3,TRAIN> (tcgen05.mma kind::tf32 3-pass forward GEMMs and kind::f16 bf16-pair adjoint GEMMs with "
                       "TMEM operands and accumulators, kind::f16 bf16-pair MN-major weight-gradient MMAs, FFMA2 tanh-jet epilogue warps)")
        try:
            with open(os.path.join(ROOT, "profiles", "tf32_peak_r01.json")) as fh:
                pk = json.load(fh)
            peak = pk["tf32_mma_tflops_n256"] / pk["passes_per_product"]
            peak_src = ("profiles/tf32_peak_r01.json: measured tcgen05 kind::tf32 MMA rate (tools/tc_probe.cu) / 3 passes; "
                        "MEASURED_PEAKS.json holds no TF32 figure")
        except Exception:
            pass
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_fused_tcgen05_r02.json")) as fh:
                nk = json.load(fh)
            traffic = nk["dram_bytes_per_point"] * n_local_pde
            clk = (sampler.summary()["sm_mhz"] or 1965) * 1e6
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            tiles_sm = -(-(-(-n_local_pde // 128)) // sms)             # 128-point tiles of the busiest SM
            pts_sm = tiles_sm * 128
            t_tensor = tiles_sm * (nk["mma_tf32_per_tile"] * nk["cycles_per_mma_tf32"] + nk["mma_bf16_per_tile"] * nk["cycles_per_mma_bf16"] +
                                   nk.get("mma_bf16_ts_per_tile", 0) * nk.get("cycles_per_mma_bf16_ts", 17.9)) / clk
            t_fma = pts_sm * nk["fma_pipe_warp_inst_per_point"] / 4.0 / clk
            t_issue = pts_sm * nk["warp_inst_per_point"] / 4.0 / clk
            extra = {"two_pipe": {"tensor_floor_ms": t_tensor * 1e3, "fma_floor_ms": t_fma * 1e3, "issue_floor_ms": t_issue * 1e3,
                                  "frac_of_floor": max(t_tensor, t_fma, t_issue) * 1e3 / k_ms,
                                  "tensor_pipe_active_pct_ncu": nk.get("tensor_pipe_active_pct"),
                                  "fma_pipe_active_pct_ncu": nk.get("fma_pipe_active_pct"),
                                  "issue_active_pct_ncu": nk.get("issue_active_pct"),
                                  "source": nk.get("source"),
                                  "note": "floors = work the shipped kernel issues on each pipe / that pipe's rate at the sampled SM clock: "
                                          "MMAs x measured cycles per MMA (profiles/tc_probes_r02.md); FMA-pipe and all warp instructions "
                                          "per point (ncu) at 1 per scheduler and cycle"}}
        except Exception:
            pass
    elif plan.engine == "fused_tf32x3":
        # hidden-layer GEMMs (97 % of the algorithmic flops) on the warp-level tensor path, 3 mma.sync passes per product;
        # `peak` is the chip's tensor roofline for 3xTF32 (tcgen05 rate / 3); the path the kernel actually uses
        # (mma.sync.m16n8k8, tools/mma_sync_probe.cu) tops out lower -- both fractions are reported
        bound, peak, peak_src = "tensor", 368.4, "fallback"
        kernel_name = ("fused_step_kernel<D,H,L,O,ORDER=2,TRAIN> (mma.sync GEMMs: tf32 m16n8k8 x3 forward, tf32 + bf16 m16n8k16 "
                       "correction passes in the gradient-only GEMMs; FFMA2 tanh-jet math)")
        extra = {}
        try:
            with open(os.path.join(ROOT, "profiles", "tf32_peak_r01.json")) as fh:
                pk = json.load(fh)
            peak = pk["tf32_mma_tflops_n256"] / pk["passes_per_product"]
            peak_src = ("profiles/tf32_peak_r01.json: measured tcgen05 kind::tf32 MMA rate (tools/tc_probe.cu) / 3 passes; "
                        "MEASURED_PEAKS.json holds no TF32 figure")
            with open(os.path.join(ROOT, "profiles", "mma_sync_peak_r01.json")) as fh:
                wk = json.load(fh)
            with open(os.path.join(ROOT, "profiles", "fp32_peak_r01.json")) as fh:
                fk = json.load(fh)
            extra = {"warp_level_path_peak": wk["tf32_m16n8k8_tflops"] / 3.0,
                     "frac_of_warp_level_path": achieved / (wk["tf32_m16n8k8_tflops"] / 3.0),
                     "fp32_fma_peak": max(fk["ffma_chain_tflops"], fk["ffma2_chain_tflops"]),
                     "vs_fp32_fma_peak": achieved / max(fk["ffma_chain_tflops"], fk["ffma2_chain_tflops"])}
        except Exception:
            pass
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic_r01.json")) as fh:
                traffic = json.load(fh)["dram_bytes_per_point"] * n_local_pde
        except Exception:
            pass
    else:
        bound, peak, peak_src = "fp32_fma", 71.7, "fallback"
        kernel_name = "fused_step_kernel<D,H,L,O,ORDER=2,TRAIN>" if plan.engine == "fused_fp32" else "layered_fp32 collocation pipeline"
        try:
            with open(os.path.join(ROOT, "profiles", "fp32_peak_r01.json")) as fh:
                pk = json.load(fh)
            peak, peak_src = max(pk["ffma_chain_tflops"], pk["ffma2_chain_tflops"]), "profiles/fp32_peak_r01.json (tools/fp32_peak.cu on this pool)"
        except Exception:
            pass
        if plan.engine == "fused_fp32":
            try:
                with open(os.path.join(ROOT, "profiles", "ncu_traffic_r01.json")) as fh:
                    tr = json.load(fh)
                traffic = tr["dram_bytes_per_point"] * n_local_pde
            except Exception:
                pass

    # ---- e2e: host-resident inputs, H2D every step, D2H of the per-term sums every step ----------
    h2d = plan.pin_host_inputs()
    T = max(1, pb.compiled.n_out_terms)
    host_out = [torch.empty(T, dtype=torch.float32, pin_memory=True) for _ in range(2)]
    read_evt = [torch.cuda.Event(), torch.cuda.Event()]
    for _ in range(3):
        plan.prefetch_inputs(); plan.commit_inputs(); s = pb.training_step(opt); host_out[0].copy_(s[:T], non_blocking=True); torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.active.set()
    e0.record()
    plan.prefetch_inputs()                             # step 0: one pinned-arena copy (every point / target array of the step)
    loss_log = 0.0
    for i in range(args.steps):
        plan.commit_inputs()                           # this step's inputs become current (device-to-device from staging)
        if i + 1 < args.steps:
            plan.prefetch_inputs()                     # the next step's inputs travel on a side stream while this step computes
        s = pb.training_step(opt)
        host_out[i & 1].copy_(s[:T], non_blocking=True)
        read_evt[i & 1].record()
        if i > 0:                                      # the caller reads the loss of EVERY step, one step behind the device:
            read_evt[(i - 1) & 1].synchronize()        # step i is already enqueued when the host waits for step i-1's sums
            loss_log += float(host_out[(i - 1) & 1][0])
    read_evt[(args.steps - 1) & 1].synchronize()
    loss_log += float(host_out[(args.steps - 1) & 1][0])
    e1.record()
    barrier()
    sampler.active.clear()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = float(t.item())
    e2e_value = n_global_pde * args.steps / (ms_e2e * 1e-3)
    if pb.allreduce_timeouts():
        raise RuntimeError("the peer-memory all-reduce timed out: the measurement is void")
    allreduce_path = ("one-shot peer-memory kernel (pinn_p2p_*)" if pb._p2p is not None else
                      "ncclAllReduce (own communicator)" if getattr(pb, "_comm", None) is not None else "torch.distributed") if world > 1 else "none"
    sampler.stop()
    # every rank samples its own GPU: a weak-scaling step waits for the slowest one, so the line carries all of them
    clocks = sampler.summary()
    if world > 1:
        mine = torch.tensor([float(clocks["sm_mhz"] or 0), float(sum(1 << i for i, r in enumerate(
            ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")) if r in clocks["reasons"]))],
            dtype=torch.float64, device=dev)
        allc = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allc, mine)
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        clocks["per_rank_sm_mhz"] = [int(c[0].item()) for c in allc]
        clocks["reasons"] = sorted({n for c in allc for i, n in enumerate(names) if int(c[1].item()) >> i & 1})
        clocks["sm_mhz_min_over_ranks"] = min(clocks["per_rank_sm_mhz"])

    # ---- the other BASELINE.json configs and strong scaling: device-timed steps, a few iterations each -----------------
    def time_config(config, total_pde, steps, warm):
        dat = problems.build_baseline_config(config, seed=1, PDE=total_pde) if total_pde else problems.build_baseline_config(config, seed=1)
        mdl = ns.TanhMLP(dat.dim, dat.hidden, dat.out_dim, device=dev, seed=3)
        ls, lt = loss_tables.build_loss_table(dat, faithful=dat.name.startswith("cavity"))
        pbx = ns.OptimizationProblem(mdl.variables, ls, lt)
        optx = ns.Adam(learning_rate=1e-2)
        for _ in range(warm):
            pbx.training_step(optx)
        extra_w = 0
        while getattr(pbx, "_graph", None) is None and pbx._graph_eligible(optx) and extra_w < 8:
            pbx.training_step(optx)
            extra_w += 1
        barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in evs:
            flush.fill_(1.0)
            a.record()
            pbx.training_step(optx)
            b.record()
        barrier()
        tt = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        n_pde = sum(cs.pointset.n for cs in pbx.compiled.sets if cs.pointset.name.upper() == "PDE") or 1
        ms = float(tt.item()) / steps
        res = {"config": config, "engine": pbx.plan.engine, "collocation_points_total": int(n_pde), "ms_per_step": ms,
               "value": n_pde / (ms * 1e-3), "unit": "pts/s", "mlp": "-".join(str(v) for v in (dat.dim, f"{dat.hidden[0]}x{len(dat.hidden)}", dat.out_dim))}
        del pbx, mdl
        torch.cuda.empty_cache()
        return res

    others, strong = [], None
    if not args.no_other_configs:
        for cfg_name in ("Poisson_Problem", "Poiseuille_Flow", "Colliding_Flow", "Cavity_Steady", "Cavity_Unsteady"):
            if cfg_name == args.config:
                continue
            base = problems.BASELINE_CONFIGS[cfg_name].get("PDE", 0)
            try:
                others.append(time_config(cfg_name, base, steps=20 if cfg_name != "Cavity_Unsteady" else 3, warm=5 if cfg_name != "Cavity_Unsteady" else 2))
            except Exception as exc:          # a config that cannot run here is reported, not hidden
                others.append({"config": cfg_name, "error": repr(exc)[:200]})
        if world > 1:
            strong = time_config(args.config, per_gpu, steps=20, warm=5)      # the N = 1 workload sharded over N GPUs
            strong["note"] = f"strong scaling: {per_gpu} collocation points in total over {world} GPUs"

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, ms, cores, sample, n_cpu = cpu_reference_step_rate(args.config, steps=3, warmup=1, budget_s=25.0, full_pde=per_gpu)
        cpu = {"value": v, "unit": "pts/s", "cores": cores, "kind": "port", "sample": sample, "ms_per_step": ms,
               "sample_points": n_cpu, "full_workload": n_cpu == per_gpu}

    if rank == 0:
        line = {
            "metric": "collocation_points_per_second_per_training_step", "value": value, "unit": "pts/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (3xtf32 tensor-core products, fp32 accumulate)" if plan.engine.endswith("tf32x3") or plan.engine == "fused_tcgen05" else "f32", "data": "synthetic",
            "config": {"workload": workload_text(args.config, per_gpu, n_global_pde, data, d, H, L, O, pb.compiled.n_params, len(losses)),
                       "engine": plan.engine, "parallelism": f"dp{world} (points sharded, SUM all-reduce of {pb.compiled.n_params + T} floats: {allreduce_path})",
                       "l2": "256 MiB device buffer rewritten between timed steps", "optimizer": "Adam(1e-2) update inside the step"},
            "e2e": {"value": e2e_value, "unit": "pts/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(T * 4),
                    "ms_per_step": ms_e2e / args.steps,
                    "pipeline": "every step's inputs are copied from pinned host memory inside the timed region; the copy of "
                                "step i+1 runs on a side stream during step i (two pinned arenas, one device staging buffer); every "
                                "step's per-term sums are copied to pinned host memory and read by the host one step behind the "
                                "device (step i+1 is enqueued before the host waits for step i)"},
            "gpu_launches": int(launches_per_step * args.steps),
            "roofline": {"bound": bound, "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": kernel_name, "kernel_ms": k_ms,
                         "flop_per_point": flops_per_point(d, H, L, O, C), "points_per_launch": n_local_pde,
                         "peak_source": peak_src,
                         "hbm_GBps": (n_local_pde * 4 * d) / (k_ms * 1e-3) * 1e-9, **extra},
            "cpu_baseline": cpu,
            "other_configs": others,
            "strong_scaling": strong,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    if os.environ.get("PINN_BENCH_WATCHDOG"):      # development aid: python stacks of a hung rank after N seconds
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["PINN_BENCH_WATCHDOG"]), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="Cavity_Steady")
    ap.add_argument("--pde-per-gpu", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
