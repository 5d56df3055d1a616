"""Turn one `ncu --set full` capture of fused_tc_kernel (+ the built library) into the committed evidence under profiles/:

    python tools/make_profiles.py gpurun_out/prof_r02_X.ncu-rep <points in the profiled launch> <tag, e.g. r02>

writes  profiles/ncu_fused_tcgen05_<tag>.md    key counters, stall reasons, opcode mix, per-source-line table
        profiles/ncu_fused_tcgen05_<tag>.json  what bench.py's roofline block reads (traffic, per-pipe instruction counts, MMAs per tile)
        profiles/sass_fused_tc_<tag>.txt       SASS of fused_tc_kernel<2,3,TRAIN> from the library (cuobjdump) + opcode histogram
Runs here (no GPU)."""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

rep, n_points, tag = sys.argv[1], int(sys.argv[2]), sys.argv[3]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "pinns_fluid_dynamics_b200", "lib", "libpinnstep.so")
KERNEL = "fused_tc_kernelILi2ELi3ELb1E"


def run(cmd):
    return subprocess.run(cmd, capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "raw", "--csv"]))))
hdr, units, row = raw[0], raw[1], raw[2]
m = dict(zip(hdr, row))
unit = dict(zip(hdr, units))


def val(key, default=0.0):
    try:
        return float(m[key].replace(",", ""))
    except Exception:
        return default


summary = run([sys.executable, os.path.join(root, "tools", "ncu_summary.py"), rep])
opmix = run([sys.executable, os.path.join(root, "tools", "ncu_opmix.py"), rep, "36"])
lines = run([sys.executable, os.path.join(root, "tools", "ncu_lines.py"), rep, KERNEL, "fused_tc.cuh", "45"])

# MMAs per 128-point tile from the SASS page: UTCHMMA with a shared-memory A descriptor (gdesc) = bf16 weight gradient, with a
# tensor-memory A operand (tmem) = tf32 forward GEMMs and bf16-pair adjoint GEMMs
src = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "source", "--csv"]))))
sh = src[1]
ix = {h: i for i, h in enumerate(sh)}
n_ss = n_ts = 0
pipe = collections.Counter()
FMA_OPS = {"FFMA", "FFMA2", "FMUL", "FMUL2", "FADD", "FADD2", "IMAD", "HFMA2", "HADD2", "HMUL2"}       # issue on the FMA pipes
for r in src[2:]:
    if len(r) < len(sh):
        continue
    toks = r[ix["Source"]].split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith("@") else toks[0]
    n = int(r[ix["Instructions Executed"]])
    base = op.split(".")[0]
    if base == "UTCHMMA":
        if "gdesc" in r[ix["Source"]].split("UTCHMMA")[1].split(",")[0]:
            n_ss += n
        else:
            n_ts += n
    if base in FMA_OPS:
        pipe["fma"] += n
    pipe["all"] += n
tiles = -(-n_points // 128) + 34          # + the boundary / fit tiles of the Cavity_Steady launch (4 x 1000 + 100 + 1 points)
dram = val("dram__bytes_read.sum") * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(unit.get("dram__bytes_read.sum", "byte"), 1.0) + \
    val("dram__bytes_write.sum") * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(unit.get("dram__bytes_write.sum", "byte"), 1.0)
inst = val("smsp__inst_executed.sum")
out = {
    "kernel": "fused_tc_kernel<2,3,TRAIN>",
    "points_in_profiled_launch": n_points,
    "dram_bytes_per_point": dram / n_points,
    "warp_inst_per_point": inst / n_points,
    "fma_pipe_warp_inst_per_point": pipe["fma"] * (inst / max(pipe["all"], 1)) / n_points,
    # tensor-memory-operand MMAs: 2 forward GEMMs x 60 kind::tf32 + 2 adjoint GEMMs x 30 kind::f16 (bf16 pairs) per tile -- SASS prints both as
    # UTCHMMA tmem[..], the kind sits in the instruction descriptor; shared-memory-operand MMAs: the bf16-pair weight gradient
    "mma_ts_per_tile": round(n_ts / tiles),
    "mma_tf32_per_tile": round(n_ts / tiles * 120 / 180),
    "mma_bf16_ts_per_tile": round(n_ts / tiles * 60 / 180),
    "mma_bf16_per_tile": round(n_ss / tiles),
    "cycles_per_mma_tf32": 17.9,
    "cycles_per_mma_bf16_ts": 17.9,        # TS bf16 K-major M128 N32 K16, profiles/tc_probes_r02.md
    "cycles_per_mma_bf16": 43.7,
    "tensor_pipe_active_pct": val("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", None),
    "fma_pipe_active_pct": val("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", None),
    "alu_pipe_active_pct": val("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", None),
    "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active", None),
    "kernel_us_under_ncu": val("gpu__time_duration.sum"),
    "source": f"ncu --set full --clock-control none, one launch ({os.path.basename(rep)}); cycles per MMA: profiles/tc_probes_r02.md; "
              f"FMA-pipe instructions = FFMA/FMUL/FADD (+ packed forms), IMAD, HFMA2 warp instructions of the SASS page scaled to smsp__inst_executed",
}
with open(os.path.join(root, "profiles", f"ncu_fused_tcgen05_{tag}.json"), "w") as fh:
    json.dump(out, fh, indent=1)

with open(os.path.join(root, "profiles", f"ncu_fused_tcgen05_{tag}.md"), "w") as fh:
    fh.write(f"# ncu --set full, fused_tc_kernel<2,3,TRAIN> (engine fused_tcgen05), Cavity_Steady {n_points} collocation points, 1 launch\n")
    fh.write("# command: ncu --set full --clock-control none --import-source on -k regex:fused_tc_kernel -s 3 -c 1 "
             f"python tools/quick_time.py cavity_steady {n_points}\n\n## counters\n```\n{summary}```\n")
    fh.write(f"\n## derived (profiles/ncu_fused_tcgen05_{tag}.json, read by bench.py)\n```\n{json.dumps(out, indent=1)}\n```\n")
    fh.write(f"\n## opcode mix (SASS page)\n```\n{opmix}```\n")
    fh.write(f"\n## stall samples by source line (tools/ncu_lines.py: SASS page joined with nvdisasm -gi of the profiled library)\n```\n{lines}```\n")

# SASS listing + opcode histogram of the shipped kernel
sass = run(["cuobjdump", "-sass", "-fun", "_ZN4pinn3ftc15fused_tc_kernelILi2ELi3ELb1EEEvPKfPKNS_6SegDevEiiPfiii", lib])
hist = collections.Counter()
for l in sass.split("\n"):
    mm = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
    if mm:
        hist[mm.group(1).split(".")[0]] += 1
with open(os.path.join(root, "profiles", f"sass_fused_tc_{tag}.txt"), "w") as fh:
    fh.write("# cuobjdump -sass of fused_tc_kernel<2,3,TRAIN> in pinns_fluid_dynamics_b200/lib/libpinnstep.so\n# static opcode histogram:\n")
    for op, c in hist.most_common():
        fh.write(f"#   {op:14s} {c}\n")
    for l in sass.split("\n"):                                  # drop the instruction encodings: keep `/*offset*/ instruction ;`
        l = re.sub(r"\s*/\* 0x[0-9a-f]{16} \*/\s*$", "", l)
        if l.strip():
            fh.write(l.rstrip() + "\n")
print(json.dumps(out, indent=1))
print("UTCHMMA static:", hist.get("UTCHMMA"), "HMMA static:", hist.get("HMMA"), "LDTM:", hist.get("LDTM"), "STTM:", hist.get("STTM"))
