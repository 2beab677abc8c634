O=gpurun_out
python tools/kernel_probe.py attn_fwd_t200 attn_com_blend_t200 ln_fwd_d108 > $O/r01d_tprobe.log 2>&1; cat $O/r01d_tprobe.log
ncu --set full --clock-control none --import-source on -k regex:attn_fwd_long -s 4 -c 1 -f -o $O/r01d_attn_long python tools/kernel_probe.py attn_fwd_t200 > $O/r01d_ncu_attn_long.log 2>&1
{ python tools/ncu_summary.py full $O/r01d_attn_long.ncu-rep; echo '```'; python tools/ncu_stalls.py $O/r01d_attn_long.ncu-rep 25 | tail -n +2; echo '```'; } > $O/r01d_full_attn_long.md 2>&1
rm -f $O/r01d_attn_long.ncu-rep
head -60 $O/r01d_full_attn_long.md
