"""Per source line of one kernel of an .ncu-rep: executed warp-instructions and stall samples (needs -lineinfo).
usage: python tools/ncu_lines2.py rep kernel-regex [top]"""
import csv, io, subprocess, sys, collections
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "-k", "regex:" + kre], capture_output=True, text=True).stdout
cur = None; agg = collections.OrderedDict(); text = {}
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if len(r) < 10 or r[0] in ("", "Line No"): continue
    try: n = int(r[6]); ie = int(r[7])
    except ValueError: continue
    k = (cur, int(r[0])); a = agg.get(k, (0, 0)); agg[k] = (a[0] + n, a[1] + ie); text[k] = r[1].strip()[:100]
ts = sum(v[0] for v in agg.values()) or 1; ti = sum(v[1] for v in agg.values()) or 1
print("kernel %s: samples %d, warp-instr %.1fM" % (kre, ts, ti / 1e6))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%5.1f%% inst %5.1f%% smp  %s:%d %s" % (100 * v[1] / ti, 100 * v[0] / ts, k[0][:14], k[1], text[k]))
