# ncu evidence of the final kernels (run under gpurun): thread-per-sequence chain kernels (C1 at 1 M series),
# likelihood gather (C5) and engine 3's contractions (C3); profiles/r02_c1_small_ncu_full.txt,
# r02_c5_final_ncu_full.txt, r02_c3_factor_ncu_full.txt
cd $GRAFT_REPO_ROOT
O=gpurun_out
cap() {  # name, kernel regex, skip, count, command...
  name=$1; rx=$2; skip=$3; cnt=$4; shift 4
  "$@" > $O/plain_$name.log 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c $cnt -o $O/$name -f "$@" > $O/ncu_$name.log 2>&1
  python tools/ncu_summary.py $O/$name.ncu-rep x > $O/${name}_ncu_full.txt 2>&1
  rm -f $O/$name.ncu-rep
}
cap r02_c1_small "k_chain_small_(forward|backward)" 2 2 python tools/prof_configs.py C1
cap r02_c5_final "k_jt_like" 2 2 python tools/prof_configs.py C5
cap r02_c3_factor_final "k_fac_contract" 0 80 env N=8 T=2 python tools/prof_c3.py
ls -la $O | tail -4
