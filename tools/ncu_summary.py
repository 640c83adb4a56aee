"""Summarise an .ncu-rep (read here with `ncu -i`): key raw metrics + stall samples per source line.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.md]"""
import csv, io, subprocess, sys, collections, re

rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__icc_request_hit_rate.pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_xu.sum", "smsp__inst_executed_op_shared_ld.sum",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_fmalite.sum", "smsp__inst_executed_pipe_xu.sum"]
print("| metric | value | unit |\n|---|---|---|", file=out)
name = dict(zip(hdr, vals)).get("Kernel Name", "")
for h, u, v in zip(hdr, units, vals):
    if h in keep or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")):
        print("| %s | %s | %s |" % (h, v, u), file=out)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# find header row
hi = [i for i, r in enumerate(rows) if "# Samples" in r][0]
h = rows[hi]; ix = {k: i for i, k in enumerate(h)}
tot = 0; per = collections.Counter(); inst = collections.Counter()
cur = None
for r in rows[hi + 1:]:
    if len(r) < len(h):
        continue
    s = r[ix["Source"]]
    a = r[ix["Address"]] if "Address" in ix else ""
    try:
        n = int(r[ix["# Samples"]] or 0); ie = int(r[ix["Instructions Executed"]] or 0)
    except ValueError:
        continue
    per[(a, s[:110])] += n; inst[(a, s[:110])] += ie
print("\nTop source lines / instructions by stall samples:\n", file=out)
allsum = sum(per.values())
for (a, s), n in per.most_common(45):
    print("%6.2f%%  inst=%-10d %s %s" % (100.0 * n / max(allsum, 1), inst[(a, s)], a[-6:], s), file=out)
