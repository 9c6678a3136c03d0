set -x
timeout -s KILL 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_gpu_tests13.log 2>&1; tail -3 gpurun_out/r2_gpu_tests13.log
python __graft_entry__.py smoke > gpurun_out/r2_smoke13.log 2>&1; tail -6 gpurun_out/r2_smoke13.log
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench13.json 2> gpurun_out/r2_bench13.err; cut -c1-240 gpurun_out/r2_bench13.json; tail -2 gpurun_out/r2_bench13.err
