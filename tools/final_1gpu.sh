set -x
timeout -s KILL 400 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_final2.log 2>&1; tail -3 gpurun_out/gpu_tests_final2.log
python bench.py --steps 30 --warmup 5 --max-seconds 200 > gpurun_out/bench_final2_n1.json 2> gpurun_out/bench_final2_n1.err; cut -c1-200 gpurun_out/bench_final2_n1.json
