#!/bin/bash
# Evidence pass for profiles/ (run under gpurun, one GPU).  Each command is first run plain (must exit 0), then under ncu.
#   bash tools/profile_round.sh <tag>
# Writes gpurun_out/<tag>_bench.json, <tag>_launches.csv and one <tag>_full_<kernel>.md per captured kernel (metrics +
# top stalled SASS instructions; the .ncu-rep files are summarised on the box and deleted to stay under the 64 MiB cap).
TAG=${1:-rXX}
O=gpurun_out
mkdir -p $O
python bench.py --steps 5 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err || { echo "bench failed"; tail -5 $O/${TAG}_bench.err; exit 1; }
python bench.py --steps 2 --warmup 3 --no-aux --no-cpu-baseline > $O/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-aux --no-cpu-baseline > $O/${TAG}_ncu_list.log 2>&1
cap() {   # cap <name> <kernel regex> <probe command...>
  local name=$1 rx=$2; shift 2
  "$@" > $O/${TAG}_plain_${name}.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 4 -c 1 -f -o $O/${TAG}_${name} "$@" > $O/${TAG}_ncu_${name}.log 2>&1
  # summarise on the box and drop the report: gpurun_out/ is capped at 64 MiB and one report is 10-20 MiB
  if [ -f $O/${TAG}_${name}.ncu-rep ]; then
    { python tools/ncu_summary.py full $O/${TAG}_${name}.ncu-rep | sed "s/^# ncu --set full summary of/## ${name} —/"
      echo '```'; python tools/ncu_stalls.py $O/${TAG}_${name}.ncu-rep 10 | tail -n +2; echo '```'; echo; } > $O/${TAG}_full_${name}.md 2>&1
    rm -f $O/${TAG}_${name}.ncu-rep
  fi
}
cap gemm_ffn1_fwd gemm_kernel python tools/gemm_probe.py ffn1_fwd
cap gemm_ffn2_dgrad_mask_cs gemm_kernel python tools/gemm_probe.py ffn2_dgrad_mask_cs
cap gemm_ffn2_fwd_res gemm_kernel python tools/gemm_probe.py ffn2_fwd_res
cap gemm_ffn1_wgrad gemm_kernel python tools/gemm_probe.py ffn1_wgrad
cap attn_fwd attn_fwd python tools/kernel_probe.py attn_fwd
cap attn_bwd attn_bwd python tools/kernel_probe.py attn_bwd
cap ln_fwd layernorm_fwd python tools/kernel_probe.py ln_fwd
cap ln_bwd layernorm_bwd python tools/kernel_probe.py ln_bwd
cap loss_fwd loss_rows python tools/kernel_probe.py loss_fwd
cap loss_bwd loss_rows python tools/kernel_probe.py loss_bwd
cap attn_fwd_t200 attn_fwd_long python tools/kernel_probe.py attn_fwd_t200
cap ln_fwd_d108 layernorm_fwd_narrow python tools/kernel_probe.py ln_fwd_d108
python tools/kernel_probe.py > $O/${TAG}_kernel_probe.log 2>&1
python tools/gemm_probe.py > $O/${TAG}_gemm_probe.log 2>&1
ls -la $O | grep ${TAG} | wc -l
