# C4 (D = 768) and C2 bench lines at HEAD, after the single-CTA backward was deleted (the D = 768 pair kernel is the only path)
set -x
timeout -s KILL 500 python bench.py --config C4 --steps 5 --warmup 3 --max-seconds 450 --no-cpu-baseline > gpurun_out/r2_bench16_c4_n1.json 2> gpurun_out/r2_bench16_c4_n1.err; cut -c1-260 gpurun_out/r2_bench16_c4_n1.json; tail -2 gpurun_out/r2_bench16_c4_n1.err
timeout -s KILL 300 python bench.py --config C2 --steps 300 --warmup 30 --no-cpu-baseline > gpurun_out/r2_bench16_c2_n1.json 2>/dev/null; cut -c1-260 gpurun_out/r2_bench16_c2_n1.json
