cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 900 --csv --log-file gpurun_out/launches_tmp.csv python tools/profile_target.py 2 64 > gpurun_out/prof_launch_tmp.log 2>&1; tail -n 2 gpurun_out/prof_launch_tmp.log
