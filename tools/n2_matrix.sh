run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline --max-seconds 90 "${@:2}" 2>>gpurun_out/n2.err | python -c "import sys,json; [print(sys.argv[1], round(json.loads(l)['ms_per_step'],4), 'ms', round(json.loads(l)['e2e']['ms_per_step'],4)) for l in sys.stdin if l.startswith('{')]" "$*"; }
timeout -s KILL 300 python -m pytest tests/test_gpu_dist.py tests/test_gpu_graphs.py -x -q -m gpu > gpurun_out/gpu_dist_g3.log 2>&1; tail -4 gpurun_out/gpu_dist_g3.log
MCLIP_DIRECT_NCCL=0 run 29551 --batch 8192 --no-graphs
run 29552 --batch 8192 --no-graphs
run 29553 --batch 8192
MCLIP_GRAPH_NCCL=1 run 29554 --batch 8192
run 29555
MCLIP_GRAPH_NCCL=1 run 29556
