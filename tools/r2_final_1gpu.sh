# closing 1-GPU sequence of round 2
set -x
timeout -s KILL 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_gpu_tests_final.log 2>&1; tail -4 gpurun_out/r2_gpu_tests_final.log
python __graft_entry__.py smoke > gpurun_out/r2_smoke_final.log 2>&1; tail -6 gpurun_out/r2_smoke_final.log
python tools/one_fused.py 8192 512 5 > gpurun_out/r2_one_fused_8192.log 2>&1; cat gpurun_out/r2_one_fused_8192.log
python tools/one_fused.py 16384 512 5 > gpurun_out/r2_one_fused_16384.log 2>&1; cat gpurun_out/r2_one_fused_16384.log
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final_n1.json 2> gpurun_out/r2_bench_final_n1.err; cut -c1-300 gpurun_out/r2_bench_final_n1.json; tail -2 gpurun_out/r2_bench_final_n1.err
timeout -s KILL 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_final_ref.json 2> gpurun_out/r2_bench_final_ref.err; cut -c1-300 gpurun_out/r2_bench_final_ref.json
timeout -s KILL 300 python bench.py --config C2 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2_bench_final_c2.json 2>/dev/null; cut -c1-260 gpurun_out/r2_bench_final_c2.json
