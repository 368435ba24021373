"""ncu --csv launch list (one row per metric) -> one line per launch.
    python tools/launch_table.py gpurun_out/launches.csv > profiles/launches_rXX.tsv
"""
import csv
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hi]
ix = {n: i for i, n in enumerate(h)}
d = OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(h):
        continue
    k = (int(r[0]), r[ix["Kernel Name"]].split("(")[0].replace("void ", ""), r[ix["Grid Size"]], r[ix["Block Size"]])
    d.setdefault(k, {})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
cols = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "smsp__inst_executed.sum"]
names = ["time_us", "dram_rd_MB", "dram_wr_MB", "l2_MB", "l1_MB", "warp_inst_M"]
scale = [1e-3, 1e-6, 1e-6, 1e-6, 1e-6, 1e-6]
print("\t".join(["id", "kernel", "grid", "block"] + names + ["share_%"]))
tot = sum(v.get(cols[0], 0.0) for v in d.values())
for k, v in d.items():
    vals = ["%.1f" % (v[c] * s) if c in v else "-" for c, s in zip(cols, scale)]
    print("\t".join([str(k[0]), k[1], k[2], k[3]] + vals + ["%.1f" % (100.0 * v.get(cols[0], 0.0) / tot if tot else 0.0)]))
print("# total time_us %.1f (cold-cache, serialised under ncu: compare shares, not absolutes)" % (tot * 1e-3))
