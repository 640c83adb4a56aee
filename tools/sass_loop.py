"""Size and instruction mix of the innermost loops of a kernel in an object file (backward branches).
usage: python tools/sass_loop.py obj.o kernel-substring [needle]   (needle: an opcode the loop must contain, e.g. LDS.S16)"""
import re, subprocess, sys, collections
obj, kern = sys.argv[1], sys.argv[2]
needle = sys.argv[3] if len(sys.argv) > 3 else None
sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
on = False; ins = []
for l in sass.split("\n"):
    if "Function :" in l: on = kern in l; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if on and m: ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(ins)}
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA\s+(?:`\(\.L_x_\d+\)|0x([0-9a-f]+))", t)
    if not m or not m.group(1): continue
    tgt = int(m.group(1), 16)
    if tgt < a and tgt in addr:
        body = ins[addr[tgt]:i + 1]
        if needle and sum(needle in x[1] for x in body) < 2: continue
        if len(body) < 40 or len(body) > 400: continue
        mix = collections.Counter(re.sub(r"^@!?U?P\d\s+", "", x[1]).split()[0].split(".")[0] for x in body)
        print("loop %#x..%#x: %d instructions; %s" % (tgt, a, len(body), ", ".join("%s %d" % kv for kv in mix.most_common(14))))
