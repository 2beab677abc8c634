#!/bin/bash
# Final 1-GPU evidence of a round: bench line, kernel probes, one full ncu capture of the long-window attention kernels.
TAG=${1:-rXX}
O=gpurun_out
python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err || { echo bench failed; tail -5 $O/${TAG}_bench.err; exit 1; }
python tools/kernel_probe.py > $O/${TAG}_kernel_probe.log 2>&1
for c in attn_fwd_t200 attn_com_blend_t200; do
  python tools/kernel_probe.py $c > /dev/null 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:attn_fwd_long -s 4 -c 1 -f -o $O/${TAG}_$c python tools/kernel_probe.py $c > $O/${TAG}_ncu_$c.log 2>&1
  { python tools/ncu_summary.py full $O/${TAG}_$c.ncu-rep | sed "s/^# ncu --set full summary of/## $c —/"; echo '```'; python tools/ncu_stalls.py $O/${TAG}_$c.ncu-rep 12 | tail -n +2; echo '```'; echo; } > $O/${TAG}_full_$c.md 2>&1
  rm -f $O/${TAG}_$c.ncu-rep
done
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2>/dev/null
cat $O/${TAG}_kernel_probe.log; cut -c1-300 $O/${TAG}_bench_ref.json
