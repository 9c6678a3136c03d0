run() { N=$1; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline --max-seconds 90 "${@:3}" 2>>gpurun_out/n8.err | tee gpurun_out/bench_n${N}_$2.json | python -c "import sys,json; [print(sys.argv[1], round(json.loads(l)['ms_per_step'],4), 'ms', 'e2e', round(json.loads(l)['e2e']['ms_per_step'],4), round(json.loads(l)['value']/1e6,2), 'M/s') for l in sys.stdin if l.startswith('{')]" "$*"; }
run 8 29571
timeout -s KILL 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29572 tools/trace_step.py 32768 > gpurun_out/trace_n8.log 2>&1; grep -c "us  dur" gpurun_out/trace_n8.log
NCCL_PROTO=Simple run 8 29573
