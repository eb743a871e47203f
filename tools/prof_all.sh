# ncu evidence for every configuration (run under gpurun; summaries are written as text, the
# reports themselves are deleted: gpurun_out/ may not exceed 64 MiB)
cd $GRAFT_REPO_ROOT
O=gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs-block --e2e-steps 1 --em-iters 1"
$B > $O/prof_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches_bench.csv $B > $O/prof_ncu_bench.log 2>&1
cap() {  # name, kernel regex, skip, count, command...
  name=$1; rx=$2; skip=$3; cnt=$4; shift 4
  "$@" > $O/plain_$name.log 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c $cnt -o $O/$name -f "$@" > $O/ncu_$name.log 2>&1
  python tools/ncu_summary.py $O/$name.ncu-rep x > $O/${name}_ncu_full.txt 2>&1
  rm -f $O/$name.ncu-rep
}
cap r02_c2_chain "k_chain_" 2 2 env REPS=2 python tools/prof_c2.py
cap r02_em "k_chain_(backward_team|stats)" 2 2 python tools/prof_em.py
cap r02_c4 "k_dense_gemm" 30 2 python tools/prof_configs.py C4
cap r02_c5 "k_jt_like" 2 2 python tools/prof_configs.py C5
cap r02_c1 "k_chain_(forward|backward)" 2 2 python tools/prof_configs.py C1
ls -la $O
