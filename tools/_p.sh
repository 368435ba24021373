mkdir -p gpurun_out
python tools/profile_step.py --steps 3 > gpurun_out/plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__t_bytes.sum,smsp__inst_executed.sum \
    --clock-control none -s 14 -c 7 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py --steps 3 > gpurun_out/ncu_list.log 2>&1
ncu --set full --import-source on --clock-control none -s 14 -c 7 -o gpurun_out/prof_step -f python tools/profile_step.py --steps 3 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
