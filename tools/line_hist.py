"""Executed warp instructions per CUDA source line: joins `nvdisasm -g` line info of the library with
the per-SASS counts of an ncu report (dev aid).  usage: line_hist.py REPORT LIB.so KERNEL_MANGLED [min_pct]"""
import csv, io, os, re, subprocess, sys, tempfile, collections
rep, so, kern = sys.argv[1:4]
minpct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.5
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
lines = []
for f in os.listdir(tmp):
    if f.endswith(".cubin"):
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if ".text." + kern + ":" in txt:
            seg = txt.split(".text." + kern + ":")[1]
            seg = seg.split("\n//--------------------- ")[0]
            cur = ("?", 0)
            for ln in seg.splitlines():
                m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
                if m:
                    cur = (os.path.basename(m.group(1)), int(m.group(2)))
                    continue
                m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
                if m:
                    lines.append((int(m.group(1), 16), cur, m.group(2)))
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h = rows[1]; body = rows[2:]
ia = h.index("Instructions Executed"); ismp = h.index("# Samples")
assert len(body) == len(lines), (len(body), len(lines))
tot = sum(int(r[ia]) for r in body)
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
for r, (addr, cur, ins) in zip(body, lines):
    op = ins.split()[1] if ins.startswith("@") else ins.split()[0]
    a = agg[cur]
    n = int(r[ia])
    a[0] += n
    a[1] += n if op.startswith(("DFMA", "DMUL", "DADD", "DSETP")) else 0
    a[2] += int(r[ismp]) if r[ismp].isdigit() else 0
    a[3] += 1
tots = sum(a[2] for a in agg.values())
print(f"total warp instr {tot}, samples {tots}, sass {len(lines)}")
src_cache = {}
def src(fn, ln):
    for root in ("jaxoplanet_b200/csrc", "."):
        p = os.path.join(root, fn)
        if os.path.exists(p):
            if p not in src_cache: src_cache[p] = open(p).read().splitlines()
            L = src_cache[p]
            return L[ln - 1].strip()[:90] if 0 < ln <= len(L) else ""
    return ""
for cur, a in sorted(agg.items(), key=lambda kv: (kv[0][0], kv[0][1])):
    if a[0] / tot * 100 >= minpct or a[2] / max(tots, 1) * 100 >= minpct:
        print(f"{cur[0]:18s}:{cur[1]:4d} instr {a[0]/tot*100:5.1f}% fp64 {a[1]/max(a[0],1)*100:3.0f}% stall-samples {a[2]/max(tots,1)*100:5.1f}% nsass {a[3]:4d} | {src(*cur)}")
print("---- by file")
byf = collections.defaultdict(lambda: [0, 0, 0])
for cur, a in agg.items():
    byf[cur[0]][0] += a[0]; byf[cur[0]][1] += a[1]; byf[cur[0]][2] += a[2]
for f, a in sorted(byf.items(), key=lambda kv: -kv[1][0]):
    print(f"{f:28s} instr {a[0]/tot*100:5.1f}% fp64-share {a[1]/max(a[0],1)*100:3.0f}% stall-samples {a[2]/max(tots,1)*100:5.1f}%")
if len(sys.argv) > 5:
    print("---- ranges of", sys.argv[5])
    fn = sys.argv[5]
    for rg in sys.argv[6:]:
        lo, hi = map(int, rg.split("-"))
        s = [0, 0, 0]
        for cur, a in agg.items():
            if cur[0] == fn and lo <= cur[1] <= hi:
                s[0] += a[0]; s[1] += a[1]; s[2] += a[2]
        print(f"{fn}:{lo}-{hi} instr {s[0]/tot*100:5.1f}% fp64-share {s[1]/max(s[0],1)*100:3.0f}% stall-samples {s[2]/max(tots,1)*100:5.1f}%")
