set -x
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/trace_step.py 32768 > gpurun_out/r2_trace_n2.log 2>&1; tail -60 gpurun_out/r2_trace_n2.log | cut -c1-150
MCLIP_CUDA_GRAPHS=1 timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 tools/trace_step.py 8192 > gpurun_out/r2_trace_n2_8192.log 2>&1; tail -60 gpurun_out/r2_trace_n2_8192.log | cut -c1-150
timeout -s KILL 600 python -m pytest tests/test_gpu_dist.py -q -m gpu -x > gpurun_out/r2_dist_n2c.log 2>&1; tail -3 gpurun_out/r2_dist_n2c.log
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --config C2 --steps 200 --warmup 20 --max-seconds 280 > gpurun_out/r2_bench_c2_n2c.json 2> gpurun_out/r2_bench_c2_n2c.err; grep '^{' gpurun_out/r2_bench_c2_n2c.json | cut -c1-260
