"""Turn the .ncu-rep files and the launch list of tools/final_run.sh (gpurun_out/) into the committed files under profiles/.
usage: python tools/make_profiles.py [tag]   (tag defaults to r02)"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
out_dir = os.path.join(ROOT, "gpurun_out")
prof = os.path.join(ROOT, "profiles")
scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__icc_request_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum"]


def summarise(rep, title, command):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw))); h = rows[0]; units = dict(zip(h, rows[1]))
    md = ["# %s" % title, "Command (ran plain first, exit 0): `%s`." % command,
          "`ncu --set full --clock-control none --import-source on`, read here with `ncu -i ... --page raw --csv`.  One section per launch.", ""]
    traffic = {}
    for r in rows[2:]:
        d = dict(zip(h, r)); name = d['Kernel Name'].split('(')[0].replace('void ', '')
        md += ["## %s" % name, "", "| metric | value | unit |", "|---|---|---|"]
        md += ["| %s | %s | %s |" % (k, d[k], units[k]) for k in keep if k in d]
        for k in h:
            if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('per_issue_active.ratio') and float(d[k] or 0) > 0.1:
                md.append("| stall %s (warps per issue) | %s | |" % (k[34:-23], d[k]))
        md.append("")
        traffic[name] = traffic.get(name, 0) + sum(float(d[m]) * scale[units[m]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    return md, traffic


def hot(rep, kernel, tiles):
    a = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_blocks.py"), rep, kernel, str(tiles), "16"], capture_output=True, text=True).stdout
    b = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_sass.py"), rep, kernel, "16"], capture_output=True, text=True).stdout
    return a + "\n" + b


# ---- launch list -------------------------------------------------------------------------------
rows = [r for r in csv.reader(open(os.path.join(out_dir, "%s_launches.csv" % tag))) if len(r) > 10]
hdr = rows[0]
out = ["# ncu launch list, %s final kernel set.  Command (ran plain first, exit 0):" % tag,
       "#   ncu --metrics gpu__time_duration.sum --clock-control none -k regex:\"transform_kernel|entropy_kernel|plan_chunks|count_ff|scan_groups|stuff_kernel\" -c 180 --csv python bench.py --steps 2 --warmup 3 --no-cpu --no-twin --no-configs",
       "# (-k filters out torch's synthetic-data kernels; a step = transform + entropy + plan_chunks + count_ff + scan_groups + stuff; the short launches are the e2e leg's chunks;",
       "#  per-launch times are cold-cache and serialised: compare SHARES)",
       "kernel,grid,block,duration_ns"]
tot = {}
for r in rows[1:]:
    d = dict(zip(hdr, r)); k = d['Kernel Name'].split('(')[0].replace('void ', '')
    out.append("%s,%s,%s,%s" % (k, d['Grid Size'].replace(',', ' '), d['Block Size'].replace(',', ' '), d['Metric Value']))
    tot.setdefault(k, []).append(float(d['Metric Value']))
s = sum(sum(v) for v in tot.values())
out.append("# shares: " + "; ".join("%s %.1f%%" % (k, 100 * sum(v) / s) for k, v in tot.items()))
open(os.path.join(prof, "%s_launches.csv" % tag), "w").write("\n".join(out) + "\n")

# ---- the headline step and the native twin ----------------------------------------------------------
md, tr = summarise(os.path.join(out_dir, "%s_final.ncu-rep" % tag), "%s final kernel set, BASELINE configs[1]: 256 x 1920x1080 RGB, IJG q75, 4:2:0 (one step)" % tag,
                   "python bench.py --steps 2 --warmup 3 --no-cpu --no-twin --no-configs")
open(os.path.join(prof, "%s_final_ncu_summary.md" % tag), "w").write("\n".join(md))
md2, tr2 = summarise(os.path.join(out_dir, "%s_twin444.ncu-rep" % tag), "%s final kernel set, byte-pinned native mode: 64 x 1920x1080 RGB, tje quality 2, 4:4:4 (one step)" % tag,
                     "python tools/prof_case.py --n 64 --qmode 0 --q 2 --sub 0 --steps 3")
open(os.path.join(prof, "%s_twin444_ncu_summary.md" % tag), "w").write("\n".join(md2))
p1 = sum(v for k, v in tr.items() if 'transform' in k or 'entropy' in k)
json.dump({"pass1_dram_bytes_per_launch_config2": int(p1),
           "source": "profiles/%s_final_ncu_summary.md: dram__bytes_read.sum + dram__bytes_write.sum of jg::transform_kernel<1,3> + jg::entropy_kernel, one step, 256 x 1080p q75 4:2:0 "
                     "(the coefficient plane between the two kernels is written once and read once: 2 x 1.59 GB on top of the 1.75 GB of pixels in + scan out)" % tag,
           "config2_kernels": {k: int(v) for k, v in tr.items()}, "native_twin_64x1080p_tje2_kernels": {k: int(v) for k, v in tr2.items()}},
          open(os.path.join(prof, "traffic.json"), "w"), indent=1)
hs = "# where the instructions, the stall samples and the shared-memory wavefronts go (tools/ncu_blocks.py, tools/ncu_sass.py on the .ncu-rep files)\n"
hs += "\n## 4:2:0 q75, entropy_kernel (391680 tiles)\n" + hot(os.path.join(out_dir, "%s_final.ncu-rep" % tag), "entropy", 391680)
hs += "\n## 4:2:0 q75, transform_kernel (per 2-MCU warp iteration: 1044480 of them)\n" + hot(os.path.join(out_dir, "%s_final.ncu-rep" % tag), "transform", 1044480)
hs += "\n## tje-2 4:4:4, entropy_kernel (194400 tiles)\n" + hot(os.path.join(out_dir, "%s_twin444.ncu-rep" % tag), "entropy", 194400)
hs += "\n## tje-2 4:4:4, transform_kernel (per 8-MCU warp iteration: 259200 of them)\n" + hot(os.path.join(out_dir, "%s_twin444.ncu-rep" % tag), "transform", 259200)
open(os.path.join(prof, "%s_final_hotspots.txt" % tag), "w").write(hs)

# ---- the fused round-1 kernels in the reference's own mode (captured at the start of the round, before anything changed) ----
for q in (2, 3):
    rep = os.path.join(out_dir, "%s_base_444_q%d.ncu-rep" % (tag, q))
    if os.path.exists(rep):
        md3, _ = summarise(rep, "round-1 fused kernel encode_tiles_kernel<0,3,0> in the reference's own mode: 64 x 1920x1080 RGB, tje quality %d, 4:4:4" % q,
                           "python tools/prof_case.py --n 64 --qmode 0 --q %d --sub 0 --steps 3" % q)
        open(os.path.join(prof, "%s_fused_444_q%d_ncu_summary.md" % (tag, q)), "w").write("\n".join(md3))

# ---- SASS ----------------------------------------------------------------------------------------
def sass_of(obj, start, stop):
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "imagecodecs_b200", "_build", obj)], capture_output=True, text=True).stdout
    keep_l, on = [], False
    for l in sass.split("\n"):
        if "Function :" in l: on = start in l and (stop is None or stop not in l)
        if on and not l.strip().startswith("/* 0x"): keep_l.append(l)
    return keep_l
for obj, start, name in (("kernel_1_3.o", "transform_kernel", "transform_420_3"), ("kernel_0_3.o", "transform_kernel", "transform_444_3"),
                         ("jpeg_entropy.o", "entropy_kernelILi0E", "entropy"), ("jpeg_stuff.o", "stuff_kernel", "stuff"), ("jpeg_stuff.o", "count_ff_kernel", "count_ff")):
    lines = sass_of(obj, start, None)
    open(os.path.join(prof, "%s_sass_%s.txt" % (tag, name)), "w").write("\n".join(lines))
    print(name, "SASS lines", len(lines), "FFMA", sum("FFMA" in l for l in lines), "FADD2", sum("FADD2" in l for l in lines),
          "UTMALDG", sum("UTMALDG" in l for l in lines), "SYNCS", sum("SYNCS" in l for l in lines))
print({k: int(v) for k, v in tr.items()})
