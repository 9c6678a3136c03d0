set -x
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x > gpurun_out/r2_kernels5.log 2>&1; tail -5 gpurun_out/r2_kernels5.log
python tools/one_fused.py 32768 512 5 > gpurun_out/r2_one_fused5.log 2>&1; cat gpurun_out/r2_one_fused5.log
timeout -s KILL 400 python bench.py --config C4 --steps 5 --warmup 3 --max-seconds 380 --no-cpu-baseline > gpurun_out/r2_bench_c4_n1.json 2> gpurun_out/r2_bench_c4_n1.err; cut -c1-300 gpurun_out/r2_bench_c4_n1.json; tail -3 gpurun_out/r2_bench_c4_n1.err
# (historical: MCLIP_BWD768_PAIR=0 selected the single-CTA D=768 backward, measured here at 68.5 ms/step and since deleted)
MCLIP_BWD768_PAIR=0 timeout -s KILL 400 python bench.py --config C4 --steps 5 --warmup 3 --max-seconds 380 --no-cpu-baseline --no-parity > gpurun_out/r2_bench_c4_n1_1cta.json 2> gpurun_out/r2_bench_c4_n1_1cta.err; cut -c1-300 gpurun_out/r2_bench_c4_n1_1cta.json
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --max-seconds 280 > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err; cut -c1-300 gpurun_out/r2_bench5.json
