mkdir -p gpurun_out
timeout 200 python tools/bwd_check.py > gpurun_out/bwd_check.log 2>&1; tail -2 gpurun_out/bwd_check.log
CPM_BWD_IMPL=tma timeout 200 python tools/bwd_check.py > gpurun_out/bwd_check_tma.log 2>&1; tail -1 gpurun_out/bwd_check_tma.log
timeout 200 python tools/profile_step.py --graph --steps 20 > gpurun_out/plain.log 2>&1; tail -5 gpurun_out/plain.log
timeout 900 python -m pytest tests -m gpu -x -q -k "backward or pooler or live or bench_workload or channels_last" > gpurun_out/tests.log 2>&1; tail -3 gpurun_out/tests.log
