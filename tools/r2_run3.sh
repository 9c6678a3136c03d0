set -x
timeout -s KILL 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "fused" > gpurun_out/r2_fused3.log 2>&1; tail -4 gpurun_out/r2_fused3.log
python tools/one_fused.py 32768 512 5 > gpurun_out/r2_one_fused3.log 2>&1; cat gpurun_out/r2_one_fused3.log
MCLIP_LIB_PATH=$PWD/mamba_clip_b200/libmclip_b200_prof.so MCLIP_DBG=16 python tools/one_fused.py 32768 512 1 > gpurun_out/r2_prof_fused3.log 2>&1; grep -c . gpurun_out/r2_prof_fused3.log
