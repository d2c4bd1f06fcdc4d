cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# r02 evidence: launch list of the final build (warm-cache, cold-cache) + ncu --set full of the kernels named in DESIGN.md
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r02_round_b64_final.csv python tools/profile_target.py 2 64 > gpurun_out/prof_a.log 2>&1; tail -n 1 gpurun_out/prof_a.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv3x3_halo64_kernel|conv3x3_halo128_pair_kernel|umma_gemm_pair_kernel|lg_attn_persist_kernel|k_knn_scan|k_nms_select|k_lg_ln_gelu' -s 12 -c 40 -o gpurun_out/r02_final_kernels python tools/profile_target.py 2 64 > gpurun_out/prof_b.log 2>&1; tail -n 2 gpurun_out/prof_b.log
ls -la gpurun_out/r02_final_kernels.ncu-rep
