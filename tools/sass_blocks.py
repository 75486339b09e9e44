"""Static per-loop instruction mix of a kernel in an nvdisasm -g listing (tools/sass_lean.sh).
usage: sass_blocks.py lean.dis kernel-substring [min_sts]   -- prints regions between labels that hold >= min_sts STS.128/STS.64"""
import collections
import re
import sys

dis, kname = sys.argv[1], sys.argv[2]
min_sts = int(sys.argv[3]) if len(sys.argv) > 3 else 4
lines = open(dis).read().split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l][0]
blocks, cur, label = [], [], "entry"
for l in lines[start + 1:]:
    if l.startswith("\t.section") or l.startswith(".text."):
        break
    m = re.match(r"^(\.L_x_\d+):", l)
    if m:
        blocks.append((label, cur))
        label, cur = m.group(1), []
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m:
        ins = re.sub(r"^@!?U?P\w+\s+", "", m.group(2))
        cur.append(ins)
blocks.append((label, cur))
tot = sum(len(b) for _, b in blocks)
print("kernel instructions:", tot)
# merge consecutive blocks into regions delimited by backward branches is overkill: report each block with stores
for label, b in blocks:
    ops = collections.Counter(i.split()[0] for i in b)
    sts = sum(v for k, v in ops.items() if k.startswith("STS.128") or k.startswith("STS.64"))
    if sts >= min_sts:
        fp = sum(v for k, v in ops.items() if k[:4] in ("DFMA", "DMUL", "DADD", "FFMA", "FMUL", "FADD"))
        lds = sum(v for k, v in ops.items() if k.startswith("LDS"))
        other = len(b) - fp - lds - sts
        top = ", ".join(f"{k}:{v}" for k, v in ops.most_common(12) if k[:4] not in ("DFMA", "DMUL", "FFMA", "FMUL"))
        print(f"{label:10s} n={len(b):4d} fp={fp:4d} lds={lds:3d} sts={sts:3d} other={other:4d} | {top}")
if len(sys.argv) > 4:   # list blocks [a, b) in order with their mix
    a, b = int(sys.argv[4]), int(sys.argv[5])
    for i in range(a, b):
        label, bl = blocks[i]
        ops = collections.Counter(x.split()[0] for x in bl)
        print(i, label, len(bl), dict(ops.most_common(8)), "|", bl[-1][:60] if bl else "")
else:
    idx = {label: i for i, (label, _) in enumerate(blocks)}
    print({k: idx[k] for k in idx if k in (".L_x_467", ".L_x_497", ".L_x_509", ".L_x_521")})
