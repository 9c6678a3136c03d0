# shape-selected persistent backward: parity tests, then split grid (0) vs persistent (1) at the other shapes the rule picks
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "persistent or block_grad" > gpurun_out/r2_gpu_tests18.log 2>&1; tail -3 gpurun_out/r2_gpu_tests18.log
MCLIP_BWD_PERSIST=-1 python tools/one_bwd.py 32768 8192 512 2>&1 | tail -1
for shape in "16384 16384" "32768 20480" "32768 10240" "65536 4096"; do
  for P in 0 1; do
    MCLIP_BWD_PERSIST=$P python tools/one_bwd.py $shape 512 2>&1 | tail -1
  done
done
