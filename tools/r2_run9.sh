set -x
timeout -s KILL 900 python -m pytest tests -q -m gpu > gpurun_out/r2_gpu_tests9.log 2>&1; tail -6 gpurun_out/r2_gpu_tests9.log
timeout -s KILL 400 python bench.py --config C4 --steps 5 --warmup 3 --max-seconds 380 --no-cpu-baseline > gpurun_out/r2_bench_c4_n1.json 2> gpurun_out/r2_bench_c4_n1.err; cut -c1-300 gpurun_out/r2_bench_c4_n1.json; tail -3 gpurun_out/r2_bench_c4_n1.err
timeout -s KILL 300 python bench.py --config C2 --steps 200 --warmup 20 --max-seconds 280 --no-cpu-baseline > gpurun_out/r2_bench_c2_n1.json 2> gpurun_out/r2_bench_c2_n1.err; cut -c1-300 gpurun_out/r2_bench_c2_n1.json
python tools/time_kernels.py 65536 768 > gpurun_out/r2_time_kernels_c4.log 2>&1; tail -8 gpurun_out/r2_time_kernels_c4.log
