# driver-style N=4 bench line at HEAD (the per-rank backward shape 8192 x 32768 now takes the persistent kernel)
set -x
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_n4b.json 2> gpurun_out/r2_bench_n4b.err; cut -c1-330 gpurun_out/r2_bench_n4b.json; tail -3 gpurun_out/r2_bench_n4b.err
