"""Does the host->device copy of the next step's inputs overlap the current loss step?  (e2e double buffering)
Serial: upload, step, read.  Overlapped: the upload of step i+1 goes to a staging buffer on a side stream while step i
runs; step i+1 starts with a device-to-device copy of the staging buffer into the plan's arena."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pinns_fluid_dynamics_b200 as ns
from pinns_fluid_dynamics_b200 import loss_tables, problems

data = problems.build_baseline_config("Cavity_Steady", seed=1)
model = ns.TanhMLP(data.dim, data.hidden, data.out_dim, device="cuda", seed=3)
losses, ltest = loss_tables.build_loss_table(data)
pb = ns.OptimizationProblem(model.variables, losses, ltest)
opt = ns.optimizers.Adam(learning_rate=1e-2)
plan = pb.plan
plan.pin_host_inputs()
T = pb.compiled.n_out_terms
host_out = torch.empty(T, dtype=torch.float32, pin_memory=True)
for _ in range(8):
    plan.upload_inputs(); s = pb.training_step(opt)
torch.cuda.synchronize()
K = 100
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(K):
    plan.upload_inputs(); s = pb.training_step(opt); host_out.copy_(s[:T], non_blocking=True); torch.cuda.current_stream().synchronize()
e1.record(); torch.cuda.synchronize()
print(f"serial      : {e0.elapsed_time(e1) / K:.4f} ms per step")

side = torch.cuda.Stream()
stage = torch.empty_like(plan._arena)
ready, consumed = torch.cuda.Event(), torch.cuda.Event()
main = torch.cuda.current_stream()
with torch.cuda.stream(side):
    stage.copy_(plan._pinned, non_blocking=True); ready.record(side)
e0.record()
for i in range(K):
    main.wait_event(ready)
    plan._arena.copy_(stage, non_blocking=True)           # device-to-device, ~8 MB
    consumed.record(main)
    with torch.cuda.stream(side):                         # next step's inputs while this step computes
        side.wait_event(consumed)
        stage.copy_(plan._pinned, non_blocking=True); ready.record(side)
    s = pb.training_step(opt); host_out.copy_(s[:T], non_blocking=True); main.synchronize()
e1.record(); torch.cuda.synchronize()
print(f"overlapped  : {e0.elapsed_time(e1) / K:.4f} ms per step")
