"""Summarise an .ncu-rep (read here, no GPU): key metrics + stall reasons + hottest source lines.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [out.md]"""
import csv, io, subprocess, sys, collections

rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = [
    "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "smsp__branch_targets_threads_divergent",
]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(f"## {d.get('Kernel Name','?')}", file=out)
    for k in KEYS:
        if k in d:
            print(f"- {k} = {d[k]} {units[hdr.index(k)]}", file=out)
    stalls = [(float(d[h].replace(',', '')), h) for h in hdr if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and d[h] not in ("", "n/a")]
    print("- stall reasons (warps per issue-active cycle):", file=out)
    for v, h in sorted(stalls, reverse=True)[:8]:
        print(f"    {h.split('issue_stalled_')[1].replace('_per_issue_active.ratio','')}: {v:.2f}", file=out)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if rows:
    # find header row
    for i, r in enumerate(rows):
        if "Source" in r and any("Executed" in c for c in r):
            h = r; body = rows[i + 1:]; break
    else:
        h = None
    if h:
        ci = h.index("Source")
        ce = [j for j, c in enumerate(h) if c.strip() == "# Instructions Executed" or c.strip() == "Instructions Executed"]
        cs = [j for j, c in enumerate(h) if "Warp Stall Sampling (All" in c]
        tot_e = 0; agg = collections.Counter(); ops = collections.Counter()
        for r in body:
            if len(r) <= ci: continue
            try: e = int(r[ce[0]]) if ce else 0
            except: e = 0
            tot_e += e
            op = r[ci].split()[0] if r[ci].split() else "?"
            if op.startswith("@"): op = (r[ci].split() + ["?", "?"])[1]
            ops[op.split(".")[0]] += e
        print(f"- SASS instructions executed (warp-level) total = {tot_e}", file=out)
        print("- top opcodes by executed count:", file=out)
        for op, e in ops.most_common(14):
            print(f"    {op}: {e} ({100.0*e/max(tot_e,1):.1f}%)", file=out)
