set -x
timeout -s KILL 900 python -m pytest tests -q -m gpu > gpurun_out/r2_gpu_tests8.log 2>&1; tail -6 gpurun_out/r2_gpu_tests8.log
timeout -s KILL 300 python bench.py --config C2 --steps 200 --warmup 20 --max-seconds 280 --no-cpu-baseline > gpurun_out/r2_bench_c2_n1.json 2> gpurun_out/r2_bench_c2_n1.err; cut -c1-300 gpurun_out/r2_bench_c2_n1.json; tail -3 gpurun_out/r2_bench_c2_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_c2_launches.csv python bench.py --config C2 --steps 3 --warmup 3 --no-cpu-baseline --no-parity --max-seconds 200 > gpurun_out/r2_ncu_c2.log 2>&1; grep -c small gpurun_out/r2_c2_launches.csv
