set -x
timeout -s KILL 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "preconverted or medium or golden" > gpurun_out/r2_k_n2d.log 2>&1; tail -3 gpurun_out/r2_k_n2d.log
timeout -s KILL 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_graphs.py tests/test_accum.py -q -m gpu -x > gpurun_out/r2_dist_n2d.log 2>&1; tail -3 gpurun_out/r2_dist_n2d.log
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 30 --warmup 10 --max-seconds 280 > gpurun_out/r2_bench_n2d.json 2> gpurun_out/r2_bench_n2d.err; grep '^{' gpurun_out/r2_bench_n2d.json | cut -c1-260; tail -2 gpurun_out/r2_bench_n2d.err
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tools/trace_step.py 32768 > gpurun_out/r2_trace_n2d.log 2>&1; grep "t=" gpurun_out/r2_trace_n2d.log | cut -c1-130; grep "step period" gpurun_out/r2_trace_n2d.log
