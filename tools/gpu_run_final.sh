# r02 evidence bundle (1 GPU): smoke, GPU test suite, bench lines of the four BASELINE configs + reference arm, launch list,
# ncu --set full summaries (summarised on the box: the .ncu-rep files are too large to travel)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" 2>&1 | tail -n 3 > gpurun_out/r02_smoke.log; cat gpurun_out/r02_smoke.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -n 12 > gpurun_out/r02_pytest_gpu.log; tail -n 3 gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02_bench_euroc_full_1gpu.json 2> gpurun_out/bench_euroc.err; tail -c 300 gpurun_out/bench_euroc.err
for c in kitti_50k sp_lg_512 mix_knn_10k; do timeout 600 python bench.py --config $c --steps 8 --warmup 3 --no-latency > gpurun_out/r02_bench_${c}_1gpu.json 2> gpurun_out/bench_$c.err; tail -c 300 gpurun_out/bench_$c.err; done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/bench_ref.err; tail -c 300 gpurun_out/bench_ref.err
python - <<'PY'
import json
for c in ["euroc_full","kitti_50k","sp_lg_512","mix_knn_10k"]:
    try:
        d=json.load(open("gpurun_out/r02_bench_%s_1gpu.json"%c))
        print(c, round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac"],3), {k:round(v,3) for k,v in d["stage_ms_per_round"].items()}, (d.get("latency_b1") or {}).get("p50_ms"))
    except Exception as ex: print(c,"FAILED",ex)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r02_round_b64_final.csv python tools/profile_target.py 2 64 > gpurun_out/prof_a.log 2>&1; tail -n 1 gpurun_out/prof_a.log
# the second (steady-state) round starts at launch ~170: SuperPoint + MixVPR + the first LightGlue layers at batch 64
timeout 1500 ncu --set full --clock-control none -k regex:'conv3x3_halo64_kernel|conv3x3_halo128_pair_kernel|umma_gemm_pair_kernel|umma_gemm_wres_kernel|lg_attn_persist_kernel|lg_ffn0_kernel|k_knn_scan|k_nms_select|stem_conv_kernel|umma_gemm_staged_kernel|umma_gemm_persist_kernel' -s 100 -c 96 -o /tmp/r02_final python tools/profile_target.py 2 64 > gpurun_out/prof_b.log 2>&1; tail -n 1 gpurun_out/prof_b.log
python tools/ncu_summary.py /tmp/r02_final.ncu-rep > gpurun_out/r02_kernels_ncu_selected.txt; grep -c "^==" gpurun_out/r02_kernels_ncu_selected.txt
