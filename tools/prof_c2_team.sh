cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests -m gpu -x -q -k "falls_back or chunked or factor_engine or c3_full" 2>&1 | tail -3
env REPS=2 python tools/prof_c2.py > $O/plain_r02_c2_team.log 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:k_chain_(forward|backward)_team" -s 2 -c 2 -o $O/r02_c2_team -f env REPS=2 python tools/prof_c2.py > $O/ncu_r02_c2_team.log 2>&1
python tools/ncu_summary.py $O/r02_c2_team.ncu-rep x > $O/r02_c2_chain_ncu_full.txt 2>&1
rm -f $O/r02_c2_team.ncu-rep
head -20 $O/r02_c2_chain_ncu_full.txt
