"""Static per-source-line SASS count of one kernel of the built library (no GPU, no profile needed): nvdisasm's inline-aware
line table, summed by the OUTERMOST line inside the kernel body file.  The tile loop of fused_tc_kernel is straight-line code
apart from the term loop and the barrier spins, so the static count of its lines is the per-thread dynamic count per tile --
the figure to watch while removing instructions.

    python tools/sass_lines.py <mangled kernel name substring> [body file suffix] [first line] [last line]
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

kern = sys.argv[1]
body = sys.argv[2] if len(sys.argv) > 2 else "fused_tc.cuh"
lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4]) if len(sys.argv) > 4 else 10 ** 9
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.environ.get("PINN_LIB", os.path.join(root, "pinns_fluid_dynamics_b200", "lib", "libpinnstep.so"))

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(sass) if l.startswith("_Z") and kern in l and l.rstrip().endswith(":"))
ann = re.compile(r'//## File "([^"]+)", line (\d+)(.*)')
inl = re.compile(r'inlined at "([^"]+)", line (\d+)')
ins = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(.*?);")
by_line = collections.defaultdict(lambda: [0, collections.Counter()])
ops_total = collections.Counter()
cur = None
total = 0
for l in sass[start + 1:]:
    if l.startswith("_Z") and l.rstrip().endswith(":"):
        break
    m = ann.search(l)
    if m:
        chain = [(m.group(1), int(m.group(2)))] + [(a, int(b)) for a, b in inl.findall(m.group(3))]
        outer = [c for c in chain if c[0].endswith(body)]
        cur = outer[-1][1] if outer else None
        continue
    m = ins.search(l)
    if m:
        t = m.group(2).split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        total += 1
        if cur is not None and lo <= cur <= hi:
            by_line[cur][0] += 1
            by_line[cur][1][op] += 1
            ops_total[op] += 1
src = open(os.path.join(root, "pinns_fluid_dynamics_b200", "csrc", body)).read().split("\n")
sel = sum(v[0] for v in by_line.values())
print(f"{kern}: {total} SASS instructions, {sel} on lines {lo}..{hi} of {body}")
for line, (n, ops) in sorted(by_line.items()):
    opc = " ".join(f"{o}:{c}" for o, c in ops.most_common(5))
    print(f"{line:5d} {n:5d}  {opc} | {src[line - 1].strip()[:100]}")
print("opcodes:", " ".join(f"{o}:{c}" for o, c in ops_total.most_common(30)))
