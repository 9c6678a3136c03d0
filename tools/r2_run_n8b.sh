set -x
run() { name=$1; shift; timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; grep '^{' gpurun_out/$name.json | cut -c1-200; tail -2 gpurun_out/$name.err; }
run r2b_bench_c3_n8 bench.py --gpus 8 --steps 40 --warmup 10 --max-seconds 280
run r2b_bench_c4_n8 bench.py --gpus 8 --config C4 --steps 10 --warmup 5 --max-seconds 280
run r2b_bench_c2_n8 bench.py --gpus 8 --config C2 --steps 200 --warmup 20 --max-seconds 280
