set -x
timeout -s KILL 600 python -m pytest tests/test_gpu_dist.py -q -m gpu -x > gpurun_out/r2_dist_n2e.log 2>&1; tail -3 gpurun_out/r2_dist_n2e.log
