"""Per CUDA-source-line stall samples + executed instructions from an .ncu-rep (cuda,sass view).
usage: python tools/ncu_lines.py rep [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
cur_file = None; ix = None; out = []
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No":
        ix = {}
        for i, k in enumerate(r):
            ix.setdefault(k, i)
        continue
    if ix is None or len(r) < 10 or r[0] == "":
        continue
    try:
        n = int(r[ix["# Samples"]]); ie = int(r[ix["Instructions Executed"]])
    except ValueError:
        continue
    out.append((n, ie, cur_file, r[0], r[1].strip()[:95]))
tot = sum(o[0] for o in out); toti = sum(o[1] for o in out)
print("total samples", tot, "total warp-instr", toti)
for n, ie, f, ln, s in sorted(out, key=lambda o: -o[0])[:top]:
    print("%5.1f%% smp %5.1f%% inst  %s:%-4s %s" % (100.0 * n / max(tot, 1), 100.0 * ie / max(toti, 1), f[:14], ln, s))
