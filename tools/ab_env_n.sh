#!/bin/bash
# A/B of environment switches at N GPUs on ONE box in ONE call.  usage: tools/ab_env_n.sh N "ENV=..." "ENV2=..."   ("" = defaults)
N=$1; shift
out=gpurun_out/ab_env_n${N}.log
: > $out
port=29600
for envs in "$@"; do
  for rep in 1 2; do
    port=$((port+1))
    line=$(env $envs python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --no-aux --no-sampling --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | tail -1)
    python - "$envs" "$rep" "$line" >> $out <<'PY'
import json, sys
d = json.loads(sys.argv[3])
r = d["roofline"]
print(f"{sys.argv[1] or 'defaults':40s} rep{sys.argv[2]}  {d['value']:10.0f} windows/s  {d['ms_per_step']:7.3f} ms/step  e2e {d['e2e']['value']:10.0f}  gemm {r['gemm_ms_per_step']:7.3f} ms  {r['achieved']:7.1f} TFLOP/s")
PY
  done
done
cat $out
