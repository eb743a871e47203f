cd $GRAFT_REPO_ROOT
python tools/prof_em.py > gpurun_out/plain_em_launches.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2s2_launches_em.csv python tools/prof_em.py > gpurun_out/ncu_em_launches.log 2>&1
tail -2 gpurun_out/plain_em_launches.log
