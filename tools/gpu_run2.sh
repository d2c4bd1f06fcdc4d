set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# launch list of 2 rounds at batch 64 (second round = steady state) + the B=1 calls
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r02_round_b64.csv python tools/profile_target.py 2 64 > gpurun_out/prof_launch.log 2>&1; tail -2 gpurun_out/prof_launch.log
# full captures: tcgen05 attention, NMS, fused kNN
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'lg_attn_umma_kernel|k_nms_select|k_knn_scan' -s 6 -c 6 -o gpurun_out/r02_attn_nms_knn python tools/profile_target.py 2 32 > gpurun_out/prof_full.log 2>&1; tail -2 gpurun_out/prof_full.log
timeout 600 ncu --set full --clock-control none -k regex:'umma_gemm_wres_kernel' -s 40 -c 6 -o gpurun_out/r02_wres python tools/profile_target.py 2 32 > gpurun_out/prof_wres.log 2>&1; tail -2 gpurun_out/prof_wres.log
# racecheck + memcheck on a small shape (mbarrier / TMEM kernels included)
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python tools/profile_target.py 1 2 160 224 > gpurun_out/r02_racecheck.log 2>&1; tail -5 gpurun_out/r02_racecheck.log
timeout 900 compute-sanitizer --tool memcheck python tools/profile_target.py 2 2 160 224 > gpurun_out/r02_memcheck.log 2>&1; tail -5 gpurun_out/r02_memcheck.log
ls -la gpurun_out | head -30
