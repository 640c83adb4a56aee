"""Turn gpurun_out/r01_final.ncu-rep + gpurun_out/r01_launches.csv into the committed files under profiles/.
usage: python tools/make_profiles.py [tag]   (tag defaults to r01)"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
rep = os.path.join(ROOT, "gpurun_out", "%s_final.ncu-rep" % tag)
prof = os.path.join(ROOT, "profiles")

# ---- launch list -------------------------------------------------------------------------------
rows = [r for r in csv.reader(open(os.path.join(ROOT, "gpurun_out", "%s_launches.csv" % tag))) if len(r) > 10]
hdr = rows[0]
out = ["# ncu launch list, %s final kernel set.  Command (ran plain first, exit 0):" % tag,
       "#   ncu --metrics gpu__time_duration.sum --clock-control none -k regex:\"encode_tiles|plan_chunks|stuff_kernel\" -c 60 --csv python bench.py --steps 2 --warmup 3 --no-cpu --no-twin",
       "# (-k filters out torch's synthetic-data kernels; a step = encode_tiles + plan_chunks + stuff; the short encode launches are the",
       "#  2-image parity check and the 28/29-image chunks of the e2e leg; per-launch times are cold-cache and serialised: compare SHARES)",
       "kernel,grid,block,duration_ns"]
tot = {}
for r in rows[1:]:
    d = dict(zip(hdr, r)); k = d['Kernel Name'].split('(')[0].replace('void ', '')
    out.append("%s,%s,%s,%s" % (k, d['Grid Size'].replace(',', ' '), d['Block Size'].replace(',', ' '), d['Metric Value']))
    tot.setdefault(k, []).append(float(d['Metric Value']))
s = sum(sum(v) for v in tot.values())
out.append("# shares: " + "; ".join("%s %.1f%%" % (k, 100 * sum(v) / s) for k, v in tot.items()))
open(os.path.join(prof, "%s_launches.csv" % tag), "w").write("\n".join(out) + "\n")

# ---- raw metrics ---------------------------------------------------------------------------------
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); h = rows[0]; units = dict(zip(h, rows[1]))
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__icc_request_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]
md = ["# %s final kernel set -- `ncu --set full --clock-control none --import-source on -k regex:\"encode_tiles|plan_chunks|stuff_kernel\" -s 9 -c 3`" % tag,
      "Command (ran plain first, exit 0): `python bench.py --steps 2 --warmup 3 --no-cpu --no-twin` (256 x 1920x1080 RGB, IJG q75, 4:2:0).",
      "The three launches are ONE step: pass 1 (encode), the chunk planner, pass 2 (stuffing).  Read here with `ncu -i ... --page raw --csv`.", ""]
traffic = {}
scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
for r in rows[2:]:
    d = dict(zip(h, r)); name = d['Kernel Name'].split('(')[0].replace('void ', '')
    md += ["## %s" % name, "", "| metric | value | unit |", "|---|---|---|"]
    md += ["| %s | %s | %s |" % (k, d[k], units[k]) for k in keep if k in d]
    for k in h:
        if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('per_issue_active.ratio') and float(d[k] or 0) > 0.1:
            md.append("| stall %s (warps per issue) | %s | |" % (k[34:-23], d[k]))
    md.append("")
    traffic[name] = sum(float(d[m]) * scale[units[m]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
open(os.path.join(prof, "%s_final_ncu_summary.md" % tag), "w").write("\n".join(md))
enc = [v for k, v in traffic.items() if 'encode' in k][0]
json.dump({"encode_420_3_dram_bytes_per_launch": int(enc),
           "source": "profiles/%s_final_ncu_summary.md: dram__bytes_read.sum + dram__bytes_write.sum of jg::encode_tiles_kernel<1,3>, one launch, 256 x 1080p q75 4:2:0" % tag,
           "all_kernels": {k: int(v) for k, v in traffic.items()}}, open(os.path.join(prof, "traffic.json"), "w"), indent=1)

# ---- source hot spots + SASS -----------------------------------------------------------------------
hs = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_funcs.py"), rep], capture_output=True, text=True, env=dict(os.environ, TOP="25")).stdout
open(os.path.join(prof, "%s_final_encode_source_hotspots.txt" % tag), "w").write(
    "# stall samples / executed warp-instructions per function and per source line (ncu --page source --print-source cuda,sass)\n" + hs)
for obj, name in (("kernel_1_3.o", "encode_tiles_420_3"), ("jpeg_stuff.o", "stuff")):
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "imagecodecs_b200", "_build", obj)], capture_output=True, text=True).stdout
    lines = [l for l in sass.split("\n") if not l.strip().startswith("/* 0x")]
    open(os.path.join(prof, "%s_sass_%s.txt" % (tag, name)), "w").write("\n".join(lines))
    print(name, "SASS lines", len(lines), "FFMA", sum("FFMA" in l for l in lines), "FADD/FMUL", sum(("FADD" in l or "FMUL" in l) for l in lines))
print({k: int(v) for k, v in traffic.items()})
