set -x
timeout -s KILL 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_graphs.py -q -m gpu -k "fused or graph" > gpurun_out/r2_fused12.log 2>&1; tail -3 gpurun_out/r2_fused12.log
python tools/one_fused.py 32768 512 5 > gpurun_out/r2_one_fused12.log 2>&1; cat gpurun_out/r2_one_fused12.log
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench12.json 2> gpurun_out/r2_bench12.err; cut -c1-300 gpurun_out/r2_bench12.json
