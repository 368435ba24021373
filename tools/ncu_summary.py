"""Summarise an ncu SASS source-page CSV: instruction mix by opcode, top stall reasons, hottest instructions.
   ncu -i X.ncu-rep --page source --csv --print-source sass --kernel-name regex:K --launch-skip N --launch-count 1 > f.csv
   python tools/ncu_summary.py f.csv
"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
body = []
for r in rows[hi + 1:]:
    if not r or r[0] in ("Address", "Kernel Name"):
        break                     # next launch's block
    if len(r) == len(hdr):
        body.append(r)
print(rows[0][1] if rows and len(rows[0]) > 1 else "")
tot_inst = sum(int(r[ix["Instructions Executed"]] or 0) for r in body)
tot_samp = sum(int(r[ix["# Samples"]] or 0) for r in body)
print("SASS lines", len(body), "warp instructions", tot_inst, "samples", tot_samp)
ops = Counter()
for r in body:
    op = r[ix["Source"]].split()[0] if r[ix["Source"]] else "?"
    if op.startswith("@"):
        op = r[ix["Source"]].split()[1]
    ops[op.split(".")[0]] += int(r[ix["Instructions Executed"]] or 0)
print("opcode mix:", ", ".join("%s %.1f%%" % (k, 100.0 * v / tot_inst) for k, v in ops.most_common(18)))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
sc = Counter()
for r in body:
    for s in stalls:
        sc[s] += int(r[ix[s]] or 0)
tot = sum(sc.values())
print("stalls:", ", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in sc.most_common(8)))
print("hottest instructions by samples:")
for r in sorted(body, key=lambda r: -int(r[ix["# Samples"]] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    top = sorted(stalls, key=lambda s: -int(r[ix[s]] or 0))[:2]
    print("  %6s %5.1f%%  %-70s %s" % (r[ix["# Samples"]], 100.0 * int(r[ix["# Samples"]]) / max(tot_samp, 1),
                                      r[ix["Source"]][:70], ",".join("%s=%s" % (s[6:], r[ix[s]]) for s in top)))
