"""Instruction-level summary of an ncu source page:  ncu -i X.ncu-rep --page source --csv > src.csv
    python tools/ncu_source.py src.csv [kernel-substring]
Prints opcode mix (share of executed warp-instructions and of stall samples) and how the executed
instructions split between the hot inner loop and the rest."""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        secs.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
for s in secs:
    if want not in s["name"]:
        continue
    hdr, data = s["hdr"], s["data"]
    iS, iN, iE = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    tot_i = sum(int(r[iE]) for r in data) or 1
    tot_s = sum(int(r[iN]) for r in data) or 1
    print("=" * 80)
    print(s["name"][:90])
    print(f"static instructions {len(data)}, executed warp-instructions {tot_i}, stall samples {tot_s}")
    c, cs = Counter(), Counter()
    for r in data:
        t = r[iS].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        c[op] += int(r[iE])
        cs[op] += int(r[iN])
    for op, n in c.most_common(24):
        print(f"  {op:10s} exec {n / tot_i:6.3f}  samples {cs[op] / tot_s:6.3f}")
    mx = max(int(r[iE]) for r in data)
    for lo, hi in ((0.5, 1.01), (0.1, 0.5), (0.02, 0.1), (-1, 0.02)):
        sel = [r for r in data if lo * mx < int(r[iE]) <= hi * mx]
        print(f"  exec count in ({lo},{hi}]*max: {len(sel):4d} static instrs, {sum(int(r[iE]) for r in sel) / tot_i:.3f} of "
              f"executed, {sum(int(r[iN]) for r in sel) / tot_s:.3f} of samples")
