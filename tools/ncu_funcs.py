"""Samples / executed instructions per function of jpeg_kernel.cuh from an .ncu-rep. usage: ncu_funcs.py rep [file.cuh]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
K = sys.argv[2] if len(sys.argv) > 2 else "jpeg_kernel.cuh"
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
cur = None; agg = {}; text = {}
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if len(r) < 10 or r[0] in ("", "Line No"): continue
    try: n = int(r[6]); ie = int(r[7])
    except ValueError: continue
    agg[(cur, int(r[0]))] = (n, ie); text[(cur, int(r[0]))] = r[1].strip()[:88]
ts = sum(v[0] for v in agg.values()); ti = sum(v[1] for v in agg.values())
import os
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "imagecodecs_b200", "csrc", K)
lines = open(path).read().split('\n')
marks = []
for i, l in enumerate(lines, 1):
    if re.match(r'^(JG_DEV|void encode_tiles_kernel|void stuff_kernel|void plan_chunks)', l) and '(' in l:
        marks.append((i, re.sub(r'\(.*', '', l).split()[-1]))
marks.append((len(lines) + 1, 'end'))
print("total samples %d, warp-instr (line-attributed) %.0fM" % (ts, ti / 1e6))
for (a, name), (b, _) in zip(marks, marks[1:]):
    s = sum(v[0] for (f, l), v in agg.items() if f == K and a <= l < b); i = sum(v[1] for (f, l), v in agg.items() if f == K and a <= l < b)
    if s * 200 > ts or i * 200 > ti: print("%-22s L%-4d samples %5.1f%%  instr %5.1f%%" % (name, a, 100 * s / ts, 100 * i / ti))
for f in sorted(set(k[0] for k in agg)):
    s = sum(v[0] for (ff, l), v in agg.items() if ff == f); i = sum(v[1] for (ff, l), v in agg.items() if ff == f)
    if f != K: print("file %-26s samples %5.1f%%  instr %5.1f%%" % (f, 100 * s / ts, 100 * i / ti))
print()
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(os.environ.get("TOP", "24"))]:
    print("%5.1f%% smp %5.1f%% inst %s:%d %s" % (100 * v[0] / ts, 100 * v[1] / ti, k[0][:12], k[1], text[k]))
