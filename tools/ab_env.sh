#!/bin/bash
# A/B of environment switches on ONE box in ONE call (box-to-box spread is +-4 %): each line = one bench.py run of the headline
# step only.  usage: tools/ab_env.sh "NAME=VAL ..." "NAME2=VAL2" ...   ("" = defaults)
out=gpurun_out/ab_env.log
: > $out
for envs in "$@"; do
  for rep in 1 2; do
    line=$(env $envs python bench.py --no-aux --no-sampling --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | tail -1)
    python - "$envs" "$rep" "$line" >> $out <<'PY'
import json, sys
d = json.loads(sys.argv[3])
r = d["roofline"]
print(f"{sys.argv[1] or 'defaults':40s} rep{sys.argv[2]}  {d['value']:10.0f} windows/s  {d['ms_per_step']:7.3f} ms/step  gemm {r['gemm_ms_per_step']:7.3f} ms  {r['achieved']:7.1f} TFLOP/s  clk {d['clocks']['sm_mhz']}")
PY
  done
done
cat $out
