cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lightglue_gpu.py tests/test_parity_exact_gpu.py tests/test_pipeline_gpu.py tests/test_stream_gpu.py -m gpu -x -q 2>&1 | tail -n 6
B="python bench.py --steps 5 --warmup 3 --no-latency --no-cpu-baseline"
for i in 1 2; do timeout 300 $B > gpurun_out/sweep_pk$i.json 2> gpurun_out/sweep_pk$i.err; python - <<PY
import json
d=json.load(open("gpurun_out/sweep_pk$i.json")); print("packed", round(d["value"]), {k:round(v,3) for k,v in d["stage_ms_per_round"].items()})
PY
done
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:lg_ffn0_kernel -s 20 -c 2 python tools/profile_target.py 2 64 2>&1 | grep -E "gpu__time|issue_active|tensor_cycles" | head -6
