"""Per-source-line profile of one kernel: joins the SASS page of an .ncu-rep (instructions executed, stall samples per
instruction) with nvdisasm's inline-aware line table of the library that was profiled, and sums by the OUTERMOST line inside the
kernel body (helpers like tanh2 are charged to their call site).  Runs here, no GPU needed.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep <mangled kernel name substring> [body file suffix] [top N]

The library on disk must be the build that was profiled (same SASS offsets)."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, kern = sys.argv[1], sys.argv[2]
body = sys.argv[3] if len(sys.argv) > 3 else "fused_tc.cuh"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "pinns_fluid_dynamics_b200", "lib", "libpinnstep.so")

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")

# line table of the kernel: offset -> (outermost body line, innermost "file:line")
start = next(i for i, l in enumerate(sass) if l.startswith("_Z") and kern in l and l.rstrip().endswith(":"))
table, cur_outer, cur_inner = {}, None, None
ann = re.compile(r'//## File "([^"]+)", line (\d+)(.*)')
inl = re.compile(r'inlined at "([^"]+)", line (\d+)')
ins = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(.*?);")
for l in sass[start + 1:]:
    if l.startswith("_Z") and l.rstrip().endswith(":"):
        break
    m = ann.search(l)
    if m:
        chain = [(m.group(1), int(m.group(2)))] + [(a, int(b)) for a, b in inl.findall(m.group(3))]
        cur_inner = f"{os.path.basename(chain[0][0])}:{chain[0][1]}"
        outer = [c for c in chain if c[0].endswith(body)]
        cur_outer = outer[-1][1] if outer else None
        continue
    m = ins.search(l)
    if m:
        table[int(m.group(1), 16)] = (cur_outer, cur_inner, m.group(2).split()[0] if not m.group(2).startswith("@") else m.group(2).split()[1])

raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
recs = [r for r in rows[2:] if len(r) >= len(hdr) and r[ix["Address"]]]
base = min(int(r[ix["Address"]], 16) for r in recs)
by_line = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot_i = tot_s = 0
for r in recs:
    off = int(r[ix["Address"]], 16) - base
    outer, inner, op = table.get(off, (None, None, "?"))
    n, s = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
    tot_i += n
    tot_s += s
    e = by_line[outer]
    e[0] += n
    e[1] += s
    e[2][op.split(".")[0]] += n
src = open(os.path.join(root, "pinns_fluid_dynamics_b200", "csrc", body)).read().split("\n")
print(f"{kern}: {tot_i} warp instructions, {tot_s} stall samples; by outermost line of {body} (top {top} by samples)")
print(f"{'line':>5s} {'inst %':>7s} {'samp %':>7s} {'samp/inst':>9s}  top opcodes | source")
for line, (n, s, ops) in sorted(by_line.items(), key=lambda kv: -kv[1][1])[:top]:
    text = src[line - 1].strip()[:90] if line else "(no line info)"
    opc = " ".join(f"{o}:{100 * c // max(n, 1)}" for o, c in ops.most_common(4))
    print(f"{str(line):>5s} {100 * n / tot_i:7.2f} {100 * s / tot_s:7.2f} {(s / tot_s) / max(n / tot_i, 1e-9):9.2f}  {opc} | {text}")
