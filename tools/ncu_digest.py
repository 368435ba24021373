"""Text digest of `ncu --set full` reports for profiles/:  python tools/ncu_digest.py a.ncu-rep [b.ncu-rep ...] > profiles/ncu_rNN_summary.txt
Per launch: duration, DRAM bytes (the `traffic` of bench.py's roofline object), L2/L1 bytes, instructions, IPC, occupancy,
limiters, top stall reasons and opcode mix (from the SASS source page)."""
import csv
import io
import subprocess
import sys
from collections import Counter

RAW = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
       'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
       'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
       'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size',
       'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static',
       'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
       'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct']


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def sass_summary(rep, idx):
    out = ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(idx), "--launch-count", "1"])
    rows = list(csv.reader(io.StringIO(out)))
    try:
        hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    except StopIteration:
        return ""
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0] != "Address"]
    tot = sum(int(r[ix["Instructions Executed"]] or 0) for r in body) or 1
    ops = Counter()
    for r in body:
        src = r[ix["Source"]].split()
        if not src:
            continue
        op = src[1] if src[0].startswith("@") and len(src) > 1 else src[0]
        ops[op.split(".")[0]] += int(r[ix["Instructions Executed"]] or 0)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    sc = Counter()
    for r in body:
        for s in stalls:
            sc[s] += int(r[ix[s]] or 0)
    st = sum(sc.values()) or 1
    return ("    opcode mix : " + ", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in ops.most_common(12)) + "\n" +
            "    stalls     : " + ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / st) for k, v in sc.most_common(7)) + "\n")


# --json PATH: also write {op: dram bytes per launch} for the step's main kernels (what bench.py reports as `traffic`)
JSON_OUT = None
ARGS = sys.argv[1:]
if "--json" in ARGS:
    i = ARGS.index("--json")
    JSON_OUT = ARGS[i + 1]
    ARGS = ARGS[:i] + ARGS[i + 2:]
DRAM = {}
OPKEY = (("roi_align_fwd_rows<1", "fwd7"), ("roi_align_fwd_rows<2", "fwd14"), ("roi_align_fwd_cols<1", "fwd7"), ("roi_align_fwd_cols<2", "fwd14"), ("bwd_tiles_staged<0, 7, 2>", "bwd7"),
         ("bwd_tiles_staged<0, 14, 2>", "bwd14"), ("bwd_tiles_staged<0,7,2>", "bwd7"), ("bwd_tiles_staged<0,14,2>", "bwd14"),
         ("stage_pyramid_f32", "stage"))

for rep in ARGS:
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    if len(rows) < 3:
        continue
    h = rows[0]
    print("== %s" % rep.split("/")[-1])
    for n, r in enumerate(rows[2:]):
        g = lambda k: r[h.index(k)] if k in h else "?"
        name = g("Kernel Name").split("(")[0].replace("void ", "")
        rd, wr = float(g('dram__bytes_read.sum')), float(g('dram__bytes_write.sum'))
        unit = rows[1][h.index('dram__bytes_read.sum')]
        scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(unit, 1.0)
        print("  [%d] %s  grid %s x block %s" % (n, name, g('launch__grid_size'), g('launch__block_size')))
        for pat, key in OPKEY:
            if pat in g("Kernel Name") and key not in DRAM:
                DRAM[key] = (rd + wr) * scale
                DRAM[key + "_kernel"] = name
        def mb(key, per=1.0):
            if key not in h or not r[h.index(key)].replace(".", "").replace("e", "").replace("+", "").replace("-", "").isdigit():
                return float("nan")
            u = rows[1][h.index(key)]
            return float(r[h.index(key)]) * per * {"Mbyte": 1, "Gbyte": 1e3, "Kbyte": 1e-3, "byte": 1e-6, "sector": 32e-6, "": 32e-6}.get(u, 1)
        dur_us = float(g('gpu__time_duration.sum')) * {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3, "second": 1e6, "s": 1e6}.get(
            rows[1][h.index('gpu__time_duration.sum')], 1.0)
        print("    duration %.1f us | DRAM read %.1f MB + write %.1f MB = traffic %.1f MB | L2 sectors %.0f MB | L1 sectors %.0f MB" % (
            dur_us, rd * scale / 1e6, wr * scale / 1e6, (rd + wr) * scale / 1e6,
            mb('lts__t_sectors.sum'), mb('SM_B.TriageCompute.l1tex__t_sectors.sum')))
        print("    warp instr %.1f M | IPC/SM %.2f | warps active %.0f%% | regs %s | smem dyn %s + static %s KB | CTAs/SM limits: regs %s, smem %s, warps %s" % (
            float(g('smsp__inst_executed.sum')) / 1e6, float(g('sm__inst_executed.avg.per_cycle_elapsed')),
            float(g('sm__warps_active.avg.pct_of_peak_sustained_active')), g('launch__registers_per_thread'),
            g('launch__shared_mem_per_block_dynamic'), g('launch__shared_mem_per_block_static'),
            g('launch__occupancy_limit_registers'), g('launch__occupancy_limit_shared_mem'), g('launch__occupancy_limit_warps')))
        print("    %% of peak: DRAM %.0f, L2 %.0f, L1/TEX %.0f | hit rates: L1 %.0f%%, L2 %.0f%%" % (
            float(g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')), float(g('lts__throughput.avg.pct_of_peak_sustained_elapsed')),
            float(g('l1tex__throughput.avg.pct_of_peak_sustained_elapsed')), float(g('l1tex__t_sector_hit_rate.pct')),
            float(g('lts__t_sector_hit_rate.pct'))))
        sys.stdout.write(sass_summary(rep, n))

if JSON_OUT:
    import json
    DRAM["source"] = "ncu --set full --clock-control none, one capture per kernel of tools/profile_step.py (per launch): " + ", ".join(a.split("/")[-1] for a in ARGS)
    json.dump(DRAM, open(JSON_OUT, "w"), indent=1)
