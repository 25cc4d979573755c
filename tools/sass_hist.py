"""Histogram of executed instructions over SASS address ranges of an ncu report (dev aid)."""
import csv, io, subprocess, sys
rep = sys.argv[1]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 150
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h = rows[1]; body = rows[2:]
ia = h.index("Instructions Executed"); it = h.index("Thread Instructions Executed"); isrc = h.index("Source"); ismp = h.index("# Samples")
tot = sum(int(r[ia]) for r in body if r[ia].isdigit())
print("total warp instr", tot, "n sass", len(body))
for b0 in range(0, len(body), B):
    seg = body[b0:b0 + B]
    e = sum(int(r[ia]) for r in seg); t = sum(int(r[it]) for r in seg); s = sum(int(r[ismp]) for r in seg if r[ismp].isdigit())
    ops = {}
    for r in seg:
        sp = r[isrc].split()
        op = sp[0] if not sp[0].startswith('@') else sp[1]
        ops[op.split('.')[0]] = ops.get(op.split('.')[0], 0) + int(r[ia])
    top = sorted(ops.items(), key=lambda x: -x[1])[:4]
    print(f"{b0:5d}-{b0+B:5d} exec {e/tot*100:5.1f}% thr/inst {t/max(e,1):5.1f} samples {s:6d}  {top}")
