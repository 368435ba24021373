# usage: bash tools/run_gpu_checks.sh [quick|full]   (run under gpurun)
mode=${1:-full}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests.log
tail -n 6 gpurun_out/tests.log
timeout 300 python tools/profile_step.py --graph --steps 20 > gpurun_out/plain.log 2>&1; echo "plain rc=$?"; cat gpurun_out/plain.log | tail -n 8
if [ "$mode" = "full" ]; then
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
  timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
  python tools/profile_step.py --steps 3 > gpurun_out/plain2.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__t_bytes.sum,smsp__inst_executed.sum --clock-control none -s 20 -c 10 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py --steps 3 > gpurun_out/ncu.log 2>&1
  tail -n 3 gpurun_out/smoke.log
fi
