"""The built library really contains the Blackwell instructions DESIGN.md claims (CPU only: cuobjdump on the in-tree .so).

A kernel that silently fell back to mma.sync, lost its TMA staging or its one-instruction tf32 rounding would still pass the
parity tests; this pins the instruction selection of the shipped binary."""
import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pinns_fluid_dynamics_b200", "lib", "libpinnstep.so")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


@pytest.fixture(scope="module")
def opcodes():
    if not os.path.exists(CUOBJDUMP):
        pytest.skip("cuobjdump not available")
    if not os.path.exists(LIB):
        import __graft_entry__
        __graft_entry__.build()
    sass = subprocess.run([CUOBJDUMP, "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per_kernel, cur = {}, None
    ins = re.compile(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)")
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = per_kernel.setdefault(m.group(1), collections.Counter())
            continue
        m = ins.search(line)
        if m and cur is not None:
            cur[m.group(1)] += 1
            cur[m.group(1).split(".")[0]] += 1 if "." in m.group(1) else 0
    return per_kernel


def _kernels(opcodes, needle):
    ks = {k: v for k, v in opcodes.items() if needle in k}
    assert ks, f"no kernel matching {needle} in the library"
    return ks


def test_fused_tcgen05_kernel_uses_tcgen05_tmem_and_tma(opcodes):
    for name, ops in _kernels(opcodes, "fused_tc_kernelILi2ELi3E").items():
        assert ops["UTCHMMA"] > 0, name                    # tcgen05.mma
        assert ops["HMMA"] == 0, name                      # no warp-level mma.sync left in this kernel
        assert ops["STTM"] > 0 and ops["LDTM"] > 0, name   # operands written to / accumulators read from tensor memory
        assert ops["UBLKCP"] > 0, name                     # the parameter vector arrives by a TMA bulk copy
        assert ops["F2FP.SATFINITE.TF32.F32.PACK_B"] > 0, name   # one-instruction round-to-nearest tf32 operand split
    train = _kernels(opcodes, "fused_tc_kernelILi2ELi3ELb1E")
    for name, ops in train.items():
        assert ops["FHFMA.BF16"] > 0 and ops["FHADD.BF16"] > 0, name   # bf16-pair remainder / read-back, mixed precision
        assert ops["STS.128"] >= 40, name                  # conflict-free 16-byte image stores (4 images x 10 per tile and thread)
        assert ops["STS.64"] <= 4, name                    # ... and none of the 8-byte halves of the first layout (80 per tile)


def test_layered_tensor_core_engine_uses_tcgen05(opcodes):
    layer = _kernels(opcodes, "tc8tc_layerI")               # pinn::tc::tc_layer<D, ORDER, MODE> (tc_layer1 / tc_layer1_grad are SIMT)
    assert all(ops["UTCHMMA"] > 0 for ops in layer.values())
    assert all(ops["UTCHMMA"] > 0 for ops in _kernels(opcodes, "tc_wgrad").values())
    assert any(ops["UBLKCP"] > 0 for ops in layer.values())


def test_warp_level_engine_keeps_the_tensor_path_for_three_input_networks(opcodes):
    # 3-32x3-3 (C = 6 channels) does not fit the tensor-memory layout of fused_tc_kernel and stays on mma.sync
    ks = {k: v for k, v in opcodes.items() if "fused_step_kernelILi3ELi32E" in k}
    assert ks and any(ops["HMMA"] > 0 for ops in ks.values())
