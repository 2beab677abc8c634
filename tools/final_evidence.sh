#!/bin/bash
# Final 1-GPU evidence of a round (run under gpurun, one GPU): bench line, reference arm, ncu launch list of the same command,
# full ncu captures of the kernels named on the command line (default: the two GEMM shapes VERDICT r01 asked for + the packed
# LayerNorm kernels), isolated kernel probes.  Each profiled command is first run plain.
#   bash tools/final_evidence.sh <tag> [caps...]
TAG=${1:-rXX}; shift
O=gpurun_out
mkdir -p $O
python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err || { echo bench failed; tail -5 $O/${TAG}_bench.err; exit 1; }
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2>/dev/null
python bench.py --steps 2 --warmup 3 --no-aux --no-cpu-baseline --no-sampling > $O/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-aux --no-cpu-baseline --no-sampling > $O/${TAG}_ncu_list.log 2>&1
cap() {   # cap <name> <kernel regex> <probe command...>
  local name=$1 rx=$2; shift 2
  "$@" > $O/${TAG}_plain_${name}.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 4 -c 1 -f -o $O/${TAG}_${name} "$@" > $O/${TAG}_ncu_${name}.log 2>&1
  if [ -f $O/${TAG}_${name}.ncu-rep ]; then
    { python tools/ncu_summary.py full $O/${TAG}_${name}.ncu-rep | sed "s/^# ncu --set full summary of/## ${name} —/"
      echo '```'; python tools/ncu_stalls.py $O/${TAG}_${name}.ncu-rep 10 | tail -n +2; echo '```'; echo; } > $O/${TAG}_full_${name}.md 2>&1
    rm -f $O/${TAG}_${name}.ncu-rep
  fi
}
CAPS=${@:-"gemm_outproj_fwd_res gemm_qkv_fwd ln_fwd ln_bwd"}
for c in $CAPS; do
  case $c in
    gemm_*) cap $c gemm_kernel python tools/gemm_probe.py ${c#gemm_} ;;
    ln_*) cap $c layernorm_ python tools/kernel_probe.py $c ;;
    attn_*) cap $c attn_ python tools/kernel_probe.py $c ;;
  esac
done
python tools/kernel_probe.py > $O/${TAG}_kernel_probe.log 2>&1
python tools/gemm_probe.py > $O/${TAG}_gemm_probe.log 2>&1
python tools/step_kernels.py 5 8 > $O/${TAG}_step_kernels.log 2>&1
cut -c1-400 $O/${TAG}_bench.json; cut -c1-300 $O/${TAG}_bench_ref.json; ls $O | grep ${TAG} | wc -l
