"""Cycles per phase of fused_tc_kernel (development aid).  Needs a -DPINN_TC_PROFILE build of the library:
   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC -DPINN_TC_PROFILE \
        -o tools/bin/libpinnstep_prof.so pinns_fluid_dynamics_b200/csrc/pinnstep.cu -ldl
   PINN_LIBPINNSTEP=tools/bin/libpinnstep_prof.so python tools/tc_phase_profile.py [n_points]"""
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
buf = np.zeros((160, 20, 16), dtype=np.uint64)
ptr = buf.ctypes.data_as(C.POINTER(C.c_ulonglong))
for _ in range(3):
    pb.plan.loss_and_grad(pb.flat)
lib.pinn_tc_profile_read(ptr)
pb.plan.loss_and_grad(pb.flat)
lib.pinn_tc_profile_read(ptr)
tiles = (n + 127) // 128 + 4 * 8 + 2
per_cta = tiles / 148.0
epi = buf[:148, :16, :].astype(np.float64).mean(axis=(0, 1)) / per_cta
mma = buf[:148, 16, :].astype(np.float64).mean(axis=0) / per_cta
print(f"fused_tc_kernel, {n} points, cycles per tile (mean over CTAs and epilogue warps), engine {pb.plan.engine}")
labels = {0: "tile bookkeeping", 1: "wait d_full(G2)", 2: "E2 (layer-2 jets, operand, images)", 3: "E2 wait/drain W2", 4: "wait d_full(G3)",
          5: "E3 jets + output-jet exchange", 6: "residuals", 7: "E3 adjoint, z-bar_3, images", 8: "wait d_full(GB3)", 9: "EB2 adjoint",
          10: "EB2 wait/drain W3", 11: "EB2 images + a_1 jets", 12: "wait d_full(GB2)", 13: "load a-bar_1 + layer 1 of next tile", 14: "EB1 math"}
tot = 0.0
for k in range(15):
    print(f"  epilogue {labels[k]:34s} {epi[k]:8.0f}")
    tot += epi[k]
print(f"  epilogue total per tile              {tot:8.0f}")
ml = {0: "wait d_loaded", 1: "wait a_ready k-step 1", 2: "wait a_ready k-step 2", 3: "wait a_ready k-step 3", 4: "wait a_ready k-step 4",
      5: "issue (GEMMs)", 6: "wait img_ready", 7: "issue (weight gradient)"}
for k in range(8):
    print(f"  mma warp {ml[k]:34s} {mma[k]:8.0f}")
