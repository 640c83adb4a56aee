"""Group the SASS of one kernel of an .ncu-rep into runs of equal execution count (~ basic blocks / loops) and print where the
executed instructions and the stall samples go.  usage: python tools/ncu_blocks.py rep kernel-regex tiles [top]"""
import csv, io, subprocess, sys
rep, kre, tiles = sys.argv[1], sys.argv[2], float(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv", "-k", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if 'Source' in r and '# Samples' in r][0]
h = rows[hi]; ix = {k: i for i, k in enumerate(h)}
data = [r for r in rows[hi + 1:] if len(r) >= len(h)]
groups = []
for n, r in enumerate(data):
    c = int(r[ix['Instructions Executed']]); s = int(r[ix['# Samples']]); w = int(r[ix['L1 Wavefronts Shared']])
    if groups and abs(groups[-1][2] - c) <= 0.02 * max(c, 1): groups[-1][1] = n; groups[-1][3] += c; groups[-1][4] += s; groups[-1][5] += w
    else: groups.append([n, n, c, c, s, w])
ti = sum(g[3] for g in groups); ts = sum(g[4] for g in groups); tw = sum(g[5] for g in groups)
print("kernel %s: %.1fM warp-instr = %.0f per tile, %d samples, %.1fM shared wavefronts = %.0f per tile" % (kre, ti / 1e6, ti / tiles, ts, tw / 1e6, tw / tiles))
for g in sorted(groups, key=lambda g: -g[4])[:top]:
    print("sass %4d-%4d n=%3d x%6.1f/tile instr %5.1f%% samples %5.1f%% wavefronts %5.1f%%  %s" % (
        g[0], g[1], g[1] - g[0] + 1, g[2] / tiles, 100 * g[3] / ti, 100 * g[4] / ts, 100 * g[5] / max(tw, 1), data[g[0]][ix['Source']].strip()[:44]))
