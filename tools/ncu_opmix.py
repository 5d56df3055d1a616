"""Opcode mix + stall samples + shared-memory wavefronts from `ncu --page source --csv` output."""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
cnt = collections.Counter(); samp = collections.Counter(); wf = collections.Counter(); wfi = collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    toks = r[ix["Source"]].split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") else toks[0]
    base = op.split(".")[0]
    if base in ("LDS", "STS", "LDG", "STG"):
        sz = [p for p in op.split(".") if p in ("64", "128", "U8", "U16")]
        base += "." + (sz[0] if sz else "32")
    n = int(r[ix["Instructions Executed"]]); tot += n
    cnt[base] += n; samp[base] += int(r[ix["# Samples"]])
    wf[base] += int(r[ix["L1 Wavefronts Shared"]] or 0); wfi[base] += int(r[ix["L1 Wavefronts Shared Ideal"]] or 0)
print(f"total warp-instructions {tot}")
for op, n in cnt.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 22):
    print(f"{op:12s} {n:12d} {100*n/tot:5.1f}%  samples {samp[op]:6d}  smem wavefronts {wf[op]:10d} (ideal {wfi[op]})")
