set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
tools/membw > gpurun_out/membw.log 2>&1
python tools/profile_step.py --steps 3 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__t_bytes.sum,smsp__inst_executed.sum --clock-control none -s 20 -c 10 --csv --log-file gpurun_out/launches_r1f.csv python tools/profile_step.py --steps 3 > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/tests.log gpurun_out/smoke.log; cat gpurun_out/membw.log
