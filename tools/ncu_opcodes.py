"""Summarise an `ncu --page source --csv` dump: executed warp-instructions and stall samples per SASS opcode.
usage: ncu -i X.ncu-rep --page source --csv > src.csv ; python tools/ncu_opcodes.py src.csv [units]
`units` divides the counts (e.g. number of warp-FFTs in the launch) to give per-unit figures."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
h = rows[1]
ia, isrc, iex, ismp = h.index("Address"), h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
ops, smp = collections.Counter(), collections.Counter()
for r in rows[2:]:
    if len(r) <= iex or not r[iex].isdigit():
        continue
    op = r[isrc].split()
    if not op:
        continue
    name = op[1] if op[0].startswith("@") else op[0]
    name = name.rstrip(";")
    key = name.split(".")[0] + ("." + name.split(".")[1] if name.startswith(("LD", "ST")) and "." in name else "")
    ops[key] += int(r[iex] or 0)
    smp[key] += int(r[ismp] or 0)
tot, tots = sum(ops.values()), sum(smp.values())
print(f"total executed warp-instructions {tot}  ({tot / units:.1f} per unit), stall samples {tots}")
for k, v in ops.most_common(40):
    print(f"{k:18s} {v:12d} {v / units:9.1f}/unit {100 * v / tot:6.2f}%   samples {100 * smp[k] / max(tots, 1):6.2f}%")
