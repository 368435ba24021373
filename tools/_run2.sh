mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "bench_workload or staging_cache or tma_backward" > gpurun_out/tests_new.log 2>&1; tail -5 gpurun_out/tests_new.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print("value",d["value"],"ms",d["ms_per_step"],"repeats",d["repeats_ms"])
print("resident",d["value_resident_channels_last"]["value"], d["value_resident_channels_last"]["ms_per_step"])
print("ops",{k:round(v["ms"],4) for k,v in d["roofline"]["ops"].items()})
print("nchw",d["roofline"]["nchw_pieces_ms"])
print("refgpu",d["roi_align"])
print("e2e",d["e2e"])
print("config0",d["config0"])
for k,v in d["nms"].items(): print(k, v["ms"], v.get("reference_gpu"), v.get("cpu_baseline"))
print("cpu",d.get("cpu_baseline"))
PY
