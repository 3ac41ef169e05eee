mkdir -p gpurun_out; O=gpurun_out
for kn in attn_fwd_tc_persist attn_bwd_dkdv_tc2 attn_bwd_dq_persist; do
  timeout 240 ncu --set full --import-source on --clock-control none -k regex:$kn -s 1 -c 1 -f -o $O/r02f_$kn python tools/attn_once.py > $O/ncu_$kn.log 2>&1
  ncu -i $O/r02f_$kn.ncu-rep --page raw --csv > $O/r02f_${kn}_raw.csv 2>/dev/null
  ncu -i $O/r02f_$kn.ncu-rep --page source --csv > $O/r02f_${kn}_source.csv 2>/dev/null
  ls -la $O/r02f_$kn.ncu-rep; rm -f $O/r02f_$kn.ncu-rep
done
ls -la $O | tail -12
