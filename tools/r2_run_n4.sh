set -x
run() { name=$1; shift; timeout -s KILL 200 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus 4 --steps 40 --warmup 10 --no-parity --no-cpu-baseline --max-seconds 180 > gpurun_out/$name.json 2> gpurun_out/$name.err; python - <<PY
import json
for l in open("gpurun_out/$name.json"):
    if l.startswith("{"):
        d = json.loads(l); print("$name", "ms/step %.4f" % d["ms_per_step"], "e2e %.4f" % d["e2e"]["ms_per_step"], "kernel_ms %.4f" % d["roofline"]["kernel_ms"])
PY
}
run r2_m4_default MCLIP_F16_PRECOPY=0
run r2_m4_precopy MCLIP_F16_PRECOPY=1
run r2_m4_default2 MCLIP_F16_PRECOPY=0
run r2_m4_precopy2 MCLIP_F16_PRECOPY=1
timeout -s KILL 600 python -m pytest tests/test_gpu_dist.py -q -m gpu -x > gpurun_out/r2_dist_n4.log 2>&1; tail -3 gpurun_out/r2_dist_n4.log
