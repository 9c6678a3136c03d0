# validation after the source split / single-CTA backward removal: tests (incl. persistent option + D=768), smoke, C3 bench
set -x
timeout -s KILL 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_gpu_tests15.log 2>&1; tail -3 gpurun_out/r2_gpu_tests15.log
python __graft_entry__.py smoke > gpurun_out/r2_smoke15.log 2>&1; tail -3 gpurun_out/r2_smoke15.log
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench15_n1.json 2> gpurun_out/r2_bench15_n1.err; cut -c1-300 gpurun_out/r2_bench15_n1.json
