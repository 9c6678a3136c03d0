set -x
timeout -s KILL 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2_gpu_tests7.log 2>&1; tail -6 gpurun_out/r2_gpu_tests7.log
python __graft_entry__.py smoke > gpurun_out/r2_smoke7.log 2>&1; tail -4 gpurun_out/r2_smoke7.log
timeout -s KILL 300 python bench.py --config C2 --steps 200 --warmup 20 --max-seconds 280 --no-cpu-baseline > gpurun_out/r2_bench_c2_n1.json 2> gpurun_out/r2_bench_c2_n1.err; cut -c1-400 gpurun_out/r2_bench_c2_n1.json; tail -3 gpurun_out/r2_bench_c2_n1.err
timeout -s KILL 300 python bench.py --config C2 --batch 512 --steps 200 --warmup 20 --max-seconds 280 --no-cpu-baseline > gpurun_out/r2_bench_c2_b512_n1.json 2> gpurun_out/r2_bench_c2_b512_n1.err; cut -c1-400 gpurun_out/r2_bench_c2_b512_n1.json
MCLIP_NO_SMALL_PATH=1 timeout -s KILL 300 python bench.py --config C2 --batch 512 --steps 200 --warmup 20 --max-seconds 280 --no-cpu-baseline > gpurun_out/r2_bench_c2_b512_n1_general.json 2>/dev/null; cut -c1-400 gpurun_out/r2_bench_c2_b512_n1_general.json
timeout -s KILL 400 python tools/e2e_stage1.py --steps 10 --warmup 3 > gpurun_out/r2_e2e_stage1_n1.json 2> gpurun_out/r2_e2e_stage1_n1.err; cat gpurun_out/r2_e2e_stage1_n1.json
