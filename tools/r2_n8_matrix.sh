# W = 8 A/B of launch-path switches at C3 (device-timed ms/step from bench.py; parity and CPU legs off)
set -x
run() { name=$1; shift; timeout -s KILL 200 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus 8 --steps 40 --warmup 10 --no-parity --no-cpu-baseline --max-seconds 180 > gpurun_out/$name.json 2> gpurun_out/$name.err; python - <<PY
import json
for l in open("gpurun_out/$name.json"):
    if l.startswith("{"):
        d = json.loads(l); print("$name", "ms/step %.4f" % d["ms_per_step"], "e2e %.4f" % d["e2e"]["ms_per_step"], "kernel_ms %.4f" % d["roofline"]["kernel_ms"])
PY
}
run r2_m8_default MCLIP_DUMMY=0
run r2_m8_persist MCLIP_BWD_PERSIST=1
run r2_m8_ll128 NCCL_PROTO=LL128
run r2_m8_precopy MCLIP_F16_PRECOPY=1
run r2_m8_nographs MCLIP_CUDA_GRAPHS=0
run r2_m8_default2 MCLIP_DUMMY=0
