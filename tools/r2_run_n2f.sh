# the driver's own N=2 launch line at HEAD (default steps / warm-up), both arms
set -x
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2_bench_n2f_ref.json 2> gpurun_out/r2_bench_n2f_ref.err; cut -c1-200 gpurun_out/r2_bench_n2f_ref.json
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 > gpurun_out/r2_bench_n2f.json 2> gpurun_out/r2_bench_n2f.err; cut -c1-300 gpurun_out/r2_bench_n2f.json; tail -3 gpurun_out/r2_bench_n2f.err
