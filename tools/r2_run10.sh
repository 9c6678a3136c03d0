set -x
timeout -s KILL 900 python -m pytest tests -q -m gpu > gpurun_out/r2_gpu_tests10.log 2>&1; tail -6 gpurun_out/r2_gpu_tests10.log
