#!/bin/bash
# A/B of the gradient-allreduce knobs at N GPUs on ONE box: NCCL channel count (= SMs taken from the persistent GEMMs) and bucket size.
N=${1:-8}
O=gpurun_out
run() {  # run <tag> <env...>
  local tag=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$((RANDOM % 90 + 10)) \
      bench.py --gpus $N --steps 10 --warmup 3 --no-aux --no-cpu-baseline > $O/ab_$tag.json 2> $O/ab_$tag.err
  python - <<PY
import json
try:
    d = json.load(open("$O/ab_$tag.json")); print("$tag", round(d["value"]), round(d["ms_per_step"], 3), round(d["e2e"]["value"]))
except Exception as e:
    print("$tag failed", e)
PY
}
run default X=1
run ch4 NCCL_MAX_NCHANNELS=4
run ch4_b32 NCCL_MAX_NCHANNELS=4 IBM_BUCKET_MB=32
run default2 X=1
