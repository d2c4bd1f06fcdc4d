cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 900 --csv --log-file gpurun_out/launches_tmp.csv python tools/profile_target.py 2 64 > gpurun_out/prof_launch_tmp.log 2>&1; tail -n 1 gpurun_out/prof_launch_tmp.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'lg_attn_persist_kernel' -s 20 -c 2 -o gpurun_out/r02_attn_persist python tools/profile_target.py 2 64 > gpurun_out/prof_full.log 2>&1; tail -n 2 gpurun_out/prof_full.log
