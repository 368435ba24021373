mkdir -p gpurun_out
cat > /tmp/nmsrun.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import cpm_r_cnn_b200 as ops
from cpm_r_cnn_b200 import synthetic as sy
dev = torch.device("cuda", 0)
gen = torch.Generator().manual_seed(1000)
b, s, seg = sy.rpn_like_candidates(gen, 16, 5, 1000)
b, s, seg = b.to(dev), s.to(dev), seg.to(dev)
for _ in range(3):
    ops.batched_nms(b, s, seg, 80, 0.7, sync=False)
torch.cuda.synchronize()
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/nms_launches.csv python /tmp/nmsrun.py > gpurun_out/ncu_nms.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/nms_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
body=rows[1:]
n=len(body)//3
for r in body[-n:]:
    print(r[ki][:70], r[vi])
PY
