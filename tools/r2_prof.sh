# round-2 profiles (1 GPU): launch list of one bench step, full ncu capture of the dominant kernels, SASS summary inputs
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_ncu.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graphs --no-parity --max-seconds 200 > gpurun_out/r2_ncu_launch.log 2>&1; tail -1 gpurun_out/r2_ncu_launch.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_tn_kernel|tc_block_grad2_kernel|tc_pair_lse2_kernel" -c 5 -o gpurun_out/r2_prof python tools/prof_kernels.py > gpurun_out/r2_ncu_full.log 2>&1; tail -2 gpurun_out/r2_ncu_full.log
ncu -i gpurun_out/r2_prof.ncu-rep --page raw --csv > gpurun_out/r2_prof_raw.csv 2>/dev/null; wc -c gpurun_out/r2_prof_raw.csv gpurun_out/r2_prof.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:"tc_block_grad2_kernel" -c 1 -o gpurun_out/r2_prof768 python tools/one_bwd.py 16384 8192 768 > gpurun_out/r2_ncu_full768.log 2>&1; tail -2 gpurun_out/r2_ncu_full768.log
ncu -i gpurun_out/r2_prof768.ncu-rep --page raw --csv > gpurun_out/r2_prof768_raw.csv 2>/dev/null; wc -c gpurun_out/r2_prof768_raw.csv
