"""Per SASS instruction of one kernel of an .ncu-rep: stall samples by reason, executions, shared-memory wavefronts (actual / ideal).
usage: python tools/ncu_sass.py rep kernel-regex [top]"""
import csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv", "-k", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if 'Source' in r and '# Samples' in r][0]
h = rows[hi]; ix = {k: i for i, k in enumerate(h)}
data = [r for r in rows[hi + 1:] if len(r) >= len(h)]
reasons = [k for k in h if k.startswith('stall_') and 'Not Issued' not in k]
ts = sum(int(r[ix['# Samples']]) for r in data) or 1
ti = sum(int(r[ix['Instructions Executed']]) for r in data) or 1
wf = sum(int(r[ix['L1 Wavefronts Shared']]) for r in data); wfi = sum(int(r[ix['L1 Wavefronts Shared Ideal']]) for r in data)
print("kernel %s: %d SASS instrs, samples %d, warp-instr %.1fM, shared wavefronts %.1fM (ideal %.1fM)" % (kre, len(data), ts, ti / 1e6, wf / 1e6, wfi / 1e6))
tot = {k: sum(int(r[ix[k]]) for r in data) for k in reasons}
print("stall totals: " + ", ".join("%s %.1f%%" % (k[6:], 100 * v / ts) for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v * 100 > ts))
print("-- top instructions by samples")
for n, r in sorted(enumerate(data), key=lambda nr: -int(nr[1][ix['# Samples']]))[:top]:
    s = int(r[ix['# Samples']])
    rs = sorted(((int(r[ix[k]]), k[6:]) for k in reasons), reverse=True)[:2]
    print("%4d %5.1f%% x%-9s %-58s %s" % (n, 100 * s / ts, r[ix['Instructions Executed']], r[ix['Source']].strip()[:58], " ".join("%s=%d" % (k, v) for v, k in rs if v)))
print("-- shared-memory instructions with excess wavefronts")
for n, r in sorted(enumerate(data), key=lambda nr: -int(nr[1][ix['L1 Wavefronts Shared']]))[:20]:
    if int(r[ix['L1 Wavefronts Shared']]) == 0: break
    print("%4d wavefronts %9s ideal %9s x%-9s %s" % (n, r[ix['L1 Wavefronts Shared']], r[ix['L1 Wavefronts Shared Ideal']], r[ix['Instructions Executed']], r[ix['Source']].strip()[:60]))
