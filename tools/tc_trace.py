"""Event trace of fused_tc_kernel (development aid): clock64 stamps of CTA 0 over tiles 2..9 of its sequence, every epilogue
warp and the MMA warp, printed relative to the tile's first event.  Needs the -DPINN_TC_PROFILE build (see tc_phase_profile.py):
   PINN_LIBPINNSTEP=tools/bin/libpinnstep_prof.so python tools/tc_trace.py [n_points] [tile 0..7]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import pinns_fluid_dynamics_b200 as ns  # noqa: E402
from pinns_fluid_dynamics_b200 import _capi, loss_tables, problems  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
data = problems.cavity_steady(seed=1, PDE=n, BC=1000, Vel=100, Pres=1, Test=1000, noise_bnd=0.01, noise_fit=0.01)
model = ns.TanhMLP(data.dim, data.hidden, data.out_dim, device="cuda", seed=3)
losses, ltest = loss_tables.build_loss_table(data)
pb = ns.OptimizationProblem(model.variables, losses, ltest)
lib = _capi.load()
for _ in range(3):
    pb.plan.loss_and_grad(pb.flat)
torch.cuda.synchronize()
buf = np.zeros((8, 20, 24), dtype=np.int64)
lib.pinn_tc_trace_read(buf.ctypes.data_as(C.POINTER(C.c_longlong)))
epi = ["G2 full", "E2 arrive h0", "E2 arrive h1", "E2 end", "G3 full", "J exchanged", "residuals done", "E3adj arrive h0", "E3adj arrive h1",
       "IMG(W3) arrive", "E3adj end", "GB3 full", "EB2 arrive h0", "EB2 arrive h1", "EB2 end", "W3 drained", "IMG(W2) arrive", "GB2 full",
       "layer1 next", "EB1 end"]
mma = {0: "G2 k1", 1: "G2 k2", 2: "G2 k3", 3: "G2 k4", 4: "G2 commit", 5: "G3 k1", 6: "G3 k2", 7: "G3 k3", 8: "G3 k4", 9: "G3 commit",
       10: "GB3 k1", 11: "GB3 k2", 12: "GB3 k3", 13: "GB3 k4", 14: "GB3 commit", 15: "GB2 k1", 16: "GB2 k2", 17: "GB2 k3", 18: "GB2 k4",
       19: "GB2 commit", 21: "W3 img seen", 20: "W3 issued", 23: "W2 img seen", 22: "W2 issued"}
st = np.zeros(16, dtype=np.int64)
lib.pinn_tc_stage_read(st.ctypes.data_as(C.POINTER(C.c_longlong)))
names = ["kernel start", "segment table staged", "terms staged", "parameters landed (TMA)", "weight images, small tables, zeroing", "tensor memory allocated",
         "first tile: layer 1 done", "last tile done", "CTA row written", "kernel end"]
print("prologue / epilogue of CTA 0, thread 0 (cycles since kernel start):")
for i, nm in enumerate(names):
    print(f"  {nm:40s} {st[i] - st[0]:9d}")
for t in ([int(sys.argv[2])] if len(sys.argv) > 2 else [2, 3]):
    tr = buf[t]
    t0 = tr[:16, 0].min()
    print(f"tile {t + 2} of CTA 0: cycles since the first epilogue warp saw G2 full; epilogue warps w = 4 (2u + v) + quadrant")
    print(f"{'event':18s} {'min':>7s} {'mean':>7s} {'max':>7s}   per warp 0..15")
    for e, name in enumerate(epi):
        v = tr[:16, e] - t0
        print(f"{name:18s} {v.min():7d} {int(v.mean()):7d} {v.max():7d}   " + " ".join(f"{x:6d}" for x in v))
    print("MMA warp:")
    for e in sorted(mma, key=lambda k: tr[16, k]):
        print(f"  {mma[e]:14s} {tr[16, e] - t0:7d}")
