# ncu evidence of the second half of round 2 (run under gpurun): the thread-per-sequence chain
# kernels at 1 M x 50 and engine 3's contractions on C3; summaries as text, reports deleted
cd $GRAFT_REPO_ROOT
O=gpurun_out
cap() {  # name, kernel regex, skip, count, command...
  name=$1; rx=$2; skip=$3; cnt=$4; shift 4
  "$@" > $O/plain_$name.log 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c $cnt -o $O/$name -f "$@" > $O/ncu_$name.log 2>&1
  python tools/ncu_summary.py $O/$name.ncu-rep x > $O/${name}_ncu_full.txt 2>&1
  rm -f $O/$name.ncu-rep
}
cap r02_c1_small "k_chain_small_(forward|backward)" 2 2 python tools/prof_configs.py C1
cap r02_c3_factor "k_fac_contract" 0 75 env N=16 T=2 python tools/prof_c3.py
ls -la $O | tail -5
