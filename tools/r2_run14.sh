set -x
timeout -s KILL 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_gpu_tests14.log 2>&1; tail -3 gpurun_out/r2_gpu_tests14.log
python __graft_entry__.py smoke > gpurun_out/r2_smoke14.log 2>&1; tail -3 gpurun_out/r2_smoke14.log
timeout -s KILL 300 python bench.py --config C2 --steps 300 --warmup 30 --no-cpu-baseline > gpurun_out/r2_bench14_c2.json 2>/dev/null; cut -c1-240 gpurun_out/r2_bench14_c2.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2_c2_launches14.csv python bench.py --config C2 --batch 512 --steps 3 --warmup 3 --no-cpu-baseline --no-parity --max-seconds 200 > /dev/null 2>&1; grep "small_" gpurun_out/r2_c2_launches14.csv | tail -2 | cut -c60-400
ncu --set full --clock-control none --import-source on -k regex:"tc_pair_lse2_kernel" -c 1 -o gpurun_out/r2_prof_fwd768 python tools/time_kernels.py 16384 768 > gpurun_out/r2_ncu_fwd768.log 2>&1; ncu -i gpurun_out/r2_prof_fwd768.ncu-rep --page raw --csv > gpurun_out/r2_prof_fwd768_raw.csv 2>/dev/null; wc -c gpurun_out/r2_prof_fwd768_raw.csv
