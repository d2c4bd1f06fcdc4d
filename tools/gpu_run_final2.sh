cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" 2>&1 | tail -n 3 > gpurun_out/r02_smoke.log; cat gpurun_out/r02_smoke.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -n 12 > gpurun_out/r02_pytest_gpu.log; tail -n 3 gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02_bench_euroc_full_1gpu.json 2> gpurun_out/bench_euroc.err; tail -c 300 gpurun_out/bench_euroc.err
for c in kitti_50k sp_lg_512 mix_knn_10k; do timeout 600 python bench.py --config $c --steps 8 --warmup 3 --no-latency > gpurun_out/r02_bench_${c}_1gpu.json 2> gpurun_out/bench_$c.err; tail -c 300 gpurun_out/bench_$c.err; done
python - <<'PY'
import json
for c in ["euroc_full","kitti_50k","sp_lg_512","mix_knn_10k"]:
    try:
        d=json.load(open("gpurun_out/r02_bench_%s_1gpu.json"%c))
        print(c, round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac"],3), {k:round(v,3) for k,v in d["stage_ms_per_round"].items()}, (d.get("latency_b1") or {}).get("p50_ms"))
    except Exception as ex: print(c,"FAILED",ex)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r02_round_b64_final.csv python tools/profile_target.py 2 64 > gpurun_out/prof_a.log 2>&1; tail -n 1 gpurun_out/prof_a.log
