# Closing 1-GPU sequence of a round (run through gpurun from the repo root).
set -x
timeout -s KILL 400 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_final.log 2>&1; tail -3 gpurun_out/gpu_tests_final.log
python __graft_entry__.py smoke > gpurun_out/smoke_final.log 2>&1; tail -4 gpurun_out/smoke_final.log
python bench.py --steps 30 --warmup 5 --max-seconds 200 > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; cut -c1-200 gpurun_out/bench_final_n1.json
# launch list (cold-cache, serialised: compare shares) and a full capture of the two tensor-core kernels
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graphs --max-seconds 200 > gpurun_out/ncu_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"tc_pair_lse2_kernel|tc_block_grad2_kernel" -c 3 -o gpurun_out/prof_final python tools/prof_kernels.py > gpurun_out/ncu_full_final.log 2>&1; tail -2 gpurun_out/ncu_full_final.log
ncu -i gpurun_out/prof_final.ncu-rep --page raw --csv > gpurun_out/prof_final_raw.csv 2>/dev/null; wc -c gpurun_out/prof_final_raw.csv gpurun_out/prof_final.ncu-rep
