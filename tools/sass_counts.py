"""Per-kernel counts of the SASS mnemonics that show which data-movement / arithmetic paths a kernel uses
(cuobjdump -sass of the in-tree library; runs in the build container, no GPU needed).

    python tools/sass_counts.py > profiles/sass_r02.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cpm_r_cnn_b200", "libcpm_ops.so")
KEYS = ["UTMALDG", "UBLKCP", "UBLKPF", "LDGSTS", "SYNCS", "FFMA2", "FFMA", "STG.E.ENL2.256", "LDG.E.128", "RED", "CREDUX", "BAR"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kern, counts, total = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        kern = kern.replace("void ", "").replace("cpm::", "")
        counts[kern] = collections.Counter()
        total[kern] = 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        total[kern] += 1
        for k in KEYS:
            if op == k or op.startswith(k + ".") or (k.count(".") and op.startswith(k)):
                counts[kern][k] += 1
print("SASS mnemonic counts per kernel of cpm_r_cnn_b200/libcpm_ops.so (sm_100a), `cuobjdump -sass`")
print("UTMALDG = cp.async.bulk.tensor (TMA tensor-map load), UBLKCP = cp.async.bulk, UBLKPF = bulk L2 prefetch, LDGSTS = cp.async,")
print("SYNCS = mbarrier ops, FFMA2 = packed fp32x2 FMA, STG.E.ENL2.256 = 32-byte store, CREDUX = redux.sync; no UTC*MMA / TMEM: the path has no contraction\n")
print("%-58s %6s  %s" % ("kernel", "instr", "  ".join("%s" % k for k in KEYS)))
for k, c in counts.items():
    if k.startswith("cub::") or k.startswith("at::") or "thrust" in k:
        continue
    print("%-58s %6d  %s" % (k[:58], total[k], "  ".join("%*d" % (len(x), c.get(x, 0)) for x in KEYS)))
