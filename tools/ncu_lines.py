"""Per-CUDA-source-line instruction/stall summary from `ncu --page source --csv --print-source cuda,sass` output."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = {}
fpath = ""
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ix = {}
        for i, h in enumerate(hdr):
            ix.setdefault(h, i)
        continue
    if hdr is None or len(r) != len(hdr) or r[0] in ("-", ""):
        continue
    if r[2] != "-":      # SASS row; only source-line rows aggregate
        continue
    key = (fpath, int(r[0]), r[1])
    inst = int(r[ix["Instructions Executed"]] or 0)
    samp = int(r[ix["# Samples"]] or 0)
    a = agg.setdefault(key, [0, 0])
    a[0] += inst
    a[1] += samp
tot = sum(v[0] for v in agg.values())
tots = sum(v[1] for v in agg.values())
print("total warp instr", tot, "samples", tots)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for (f, ln, src), (inst, samp) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
    print("%-18s %4d  inst %5.1f%%  samp %5.1f%% | %s" % (f, ln, 100.0 * inst / tot, 100.0 * samp / max(tots, 1), src.strip()[:100]))
