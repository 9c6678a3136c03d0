# round 2, first GPU call: microbenchmarks, the new shared-recompute backward, headline-size parity, first bench lines
set -x
nvidia-smi -L; nproc; free -g | head -2
tools/l2_reduce_bench > gpurun_out/r2_l2_reduce.log 2>&1; tail -5 gpurun_out/r2_l2_reduce.log
timeout -s KILL 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "fused" > gpurun_out/r2_fused.log 2>&1; tail -15 gpurun_out/r2_fused.log
timeout -s KILL 900 python -m pytest tests -q -m gpu -s > gpurun_out/r2_gpu_tests.log 2>&1; tail -25 gpurun_out/r2_gpu_tests.log
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --max-seconds 280 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; cut -c1-600 gpurun_out/r2_bench1.json; tail -3 gpurun_out/r2_bench1.err
MCLIP_FUSED_BWD=0 timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --max-seconds 280 --no-cpu-baseline > gpurun_out/r2_bench1_nofuse.json 2> gpurun_out/r2_bench1_nofuse.err; cut -c1-300 gpurun_out/r2_bench1_nofuse.json
timeout -s KILL 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; cut -c1-400 gpurun_out/r2_bench_ref.json
