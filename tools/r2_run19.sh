# full GPU test suite + smoke at HEAD (shape-selected persistent backward is the default now)
set -x
timeout -s KILL 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2_gpu_tests19.log 2>&1; tail -3 gpurun_out/r2_gpu_tests19.log
python __graft_entry__.py smoke > gpurun_out/r2_smoke19.log 2>&1; tail -2 gpurun_out/r2_smoke19.log
