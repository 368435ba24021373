# usage: bash tools/run_gpu_checks.sh [quick|full|profile]   (run under gpurun; everything lands in gpurun_out/)
mode=${1:-full}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests.log
tail -n 4 gpurun_out/tests.log
timeout 300 python tools/profile_step.py --graph --steps 30 > gpurun_out/plain.log 2>&1; echo "plain rc=$?"; tail -n 8 gpurun_out/plain.log
timeout 300 python tools/profile_step.py --graph --steps 30 --resident > gpurun_out/plain_resident.log 2>&1; tail -n 6 gpurun_out/plain_resident.log
if [ "$mode" != "quick" ]; then
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
  timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
  tail -n 2 gpurun_out/smoke.log
fi
if [ "$mode" = "profile" ]; then
  # launch list of one step (step 3 of 3: 7 launches = stage, fwd7, prepare, bwd7, fwd14, prepare, bwd14), then
  # one `--set full` capture per kernel of that step; each only after the same command has exited 0 without ncu
  python tools/profile_step.py --steps 3 > gpurun_out/plain2.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__t_bytes.sum,smsp__inst_executed.sum \
      --clock-control none -s 14 -c 7 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py --steps 3 > gpurun_out/ncu_list.log 2>&1
  ncu --set full --import-source on --clock-control none -s 14 -c 7 -o gpurun_out/prof_step -f python tools/profile_step.py --steps 3 > gpurun_out/ncu_full.log 2>&1
  python tools/profile_step.py --steps 1 --nms > /dev/null 2>&1 && \
  ncu --set full --clock-control none -k regex:'nms_mask|nms_sweep' -c 6 -o gpurun_out/prof_nms -f python tools/profile_step.py --steps 1 --nms > gpurun_out/ncu_nms.log 2>&1
  ls -la gpurun_out/*.ncu-rep
fi
