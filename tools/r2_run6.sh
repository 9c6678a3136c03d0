set -x
timeout -s KILL 300 python -m pytest tests/test_gpu_next_rows.py -q -m gpu -x -s > gpurun_out/r2_next_rows6.log 2>&1; tail -8 gpurun_out/r2_next_rows6.log
timeout -s KILL 400 python tools/e2e_stage1.py --steps 10 --warmup 3 > gpurun_out/r2_e2e_stage1_n1.json 2> gpurun_out/r2_e2e_stage1_n1.err; cat gpurun_out/r2_e2e_stage1_n1.json; tail -3 gpurun_out/r2_e2e_stage1_n1.err
