# usage: bash tools/quick_bwd.sh   (under gpurun) -- RoIAlign parity tests, then the per-op step timing (both pooled layouts)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "roi_align or backward or pooler or channels_last or inference_shaped or bf16" > gpurun_out/tests_quick.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/tests_quick.log
timeout 200 python tools/profile_step.py --graph --steps 30 2>&1 | grep -E "fwd|bwd|rror" | tee gpurun_out/quick.log
timeout 200 python tools/profile_step.py --graph --steps 30 --cl 2>&1 | grep -E "fwd|bwd|rror" | tee -a gpurun_out/quick.log
timeout 200 python tools/bf16_probe.py 2>&1 | tail -3 | tee -a gpurun_out/quick.log
