"""profiles/r2_walk_eval.json (or r2_walk_items.json) from an ncu --set full report of k_walk_eval / k_walk_items
inside bench.py.  usage: python tools/make_walk_profile.py REPORT.ncu-rep ITEMS [k_walk_items]"""
import collections, csv, io, json, subprocess, sys
rep, items = sys.argv[1], int(sys.argv[2])
kern = sys.argv[3] if len(sys.argv) > 3 else "k_walk_eval"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); h, u, r = rows[0], rows[1], rows[2]; d = dict(zip(h, r)); un = dict(zip(h, u))
g = lambda k: float(d[k].replace(",", ""))
def bytes_of(k):
    v = g(k); unit = un[k].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[unit]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
for i, rr in enumerate(rows):
    if "Source" in rr and any("Executed" in c for c in rr):
        hh, body = rr, rows[i + 1:]; break
ci, cw, ct = hh.index("Source"), hh.index("Instructions Executed"), hh.index("Predicated-On Thread Instructions Executed")
thr, warp = collections.Counter(), collections.Counter()
for rr in body:
    toks = rr[ci].split() if len(rr) > ci else []
    if not toks: continue
    op = (toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]).split(".")[0]
    try: warp[op] += int(rr[cw]); thr[op] += int(rr[ct])
    except ValueError: pass
out = dict(kernel=d["Kernel Name"], workload="bench.py config 3, one GPU: 4096 sets x 1e5 times", items=items,
           dfma=thr["DFMA"], dmul=thr["DMUL"], dadd=thr["DADD"], dsetp=thr["DSETP"], warp_inst=int(g("smsp__inst_executed.sum")),
           fp64_warp_inst=sum(warp[k] for k in ("DFMA", "DMUL", "DADD", "DSETP")),
           pipe_fp64_pct=round(g("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"), 2),
           issue_active_pct=round(g("smsp__issue_active.avg.pct_of_peak_sustained_active"), 2),
           duration_us_under_ncu=g("gpu__time_duration.sum"), registers=int(g("launch__registers_per_thread")),
           grid=int(g("launch__grid_size")), dram_bytes_read=bytes_of("dram__bytes_read.sum"), dram_bytes_write=bytes_of("dram__bytes_write.sum"),
           note="predicated-on thread instructions summed over the SASS lines of the ncu source page; a launch on the same seeded "
                "inputs executes the same instructions, so bench.py divides these by the LIVE kernel duration",
           source="ncu --set full --clock-control none --import-source on -k regex:%s -s 12 -c 1 python bench.py --steps 3 "
                  "--warmup 3 --no-secondary --no-cpu (tools/gpu_bench_round.sh); summary: profiles/r2_bench_%s.md" % (kern, kern[2:]))
json.dump(out, open("profiles/r2_%s.json" % kern[2:], "w"), indent=1)
print(json.dumps(out, indent=1))
