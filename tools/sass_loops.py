"""Opcode histogram of the innermost loops of one kernel:  python tools/sass_loops.py <obj|so> <kernel substring> [min_len]
(cuobjdump -sass; a loop = a backward branch; prints the loops with the most IMAD.WIDE first)"""
import collections, re, subprocess, sys
obj, name = sys.argv[1], sys.argv[2]
min_len = int(sys.argv[3]) if len(sys.argv) > 3 else 16
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout.splitlines()
ins, on = [], False
for l in txt:
    if "Function :" in l:
        on = name in l
        continue
    if not on:
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
print(f"{len(ins)} instructions")
loops = []
for a, t in ins:
    m = re.search(r"BRA(?:\.U)? (?:!?U?P\d, )?0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        loops.append((int(m.group(1), 16), a))
def op(t):
    w = t.split()
    return w[1] if w[0].startswith("@") else w[0]
out = []
for lo, hi in loops:
    if hi - lo > 16 * 1200:
        continue  # outer loops
    body = [t for a, t in ins if lo <= a <= hi]
    if len(body) < min_len:
        continue
    c = collections.Counter(op(t) for t in body)
    out.append((c.get("IMAD.WIDE.U32", 0), lo, hi, len(body), c))
for w, lo, hi, n, c in sorted(out, key=lambda x: -x[0])[:4]:
    print(f"loop 0x{lo:x}-0x{hi:x}: {n} instr; " + ", ".join(f"{v} {k}" for k, v in c.most_common(14)))
