set -x
nvidia-smi -L | wc -l
timeout -s KILL 600 python -m pytest tests/test_gpu_dist.py -q -m gpu > gpurun_out/r2_dist_n8.log 2>&1; tail -5 gpurun_out/r2_dist_n8.log
run() { name=$1; shift; timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; grep '^{' gpurun_out/$name.json | cut -c1-260; tail -2 gpurun_out/$name.err; }
run r2_bench_c3_n8 bench.py --gpus 8 --steps 30 --warmup 10 --max-seconds 280
run r2_bench_c4_n8 bench.py --gpus 8 --config C4 --steps 10 --warmup 5 --max-seconds 280
run r2_bench_c2_n8 bench.py --gpus 8 --config C2 --steps 200 --warmup 20 --max-seconds 280
run r2_e2e_stage1_n8 tools/e2e_stage1.py --steps 10 --warmup 3
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 4 --steps 30 --warmup 10 --max-seconds 180 > gpurun_out/r2_bench_c3_n4.json 2> gpurun_out/r2_bench_c3_n4.err; grep '^{' gpurun_out/r2_bench_c3_n4.json | cut -c1-260
