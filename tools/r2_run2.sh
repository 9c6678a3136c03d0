set -x
python tools/one_fused.py 32768 512 5 > gpurun_out/r2_one_fused.log 2>&1; cat gpurun_out/r2_one_fused.log
ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_fused_launches.csv python tools/one_fused.py 32768 512 1 > gpurun_out/r2_ncu_fused.log 2>&1; tail -2 gpurun_out/r2_ncu_fused.log
