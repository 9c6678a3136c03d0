set -x
nvidia-smi -L | head -3
timeout -s KILL 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_graphs.py -q -m gpu -x > gpurun_out/r2_dist_n2.log 2>&1; tail -8 gpurun_out/r2_dist_n2.log
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --max-seconds 280 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; cut -c1-400 gpurun_out/r2_bench_n2.json; tail -3 gpurun_out/r2_bench_n2.err
