cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mixvpr_gpu.py tests/test_gemm_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -n 6
B="python bench.py --steps 5 --warmup 3 --no-latency --no-cpu-baseline --no-preflight"
run() { name=$1; shift; env "$@" timeout 300 $B > gpurun_out/sweep_$name.json 2> gpurun_out/sweep_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sweep_$name.json")); print("$name", round(d["value"]), {k:round(v,3) for k,v in d["stage_ms_per_round"].items()})
except Exception as ex: print("$name FAILED", ex)
PY
tail -n 2 gpurun_out/sweep_$name.err
}
run s2conv0 DV_MIX_S2CONV=0
run s2conv1 DV_MIX_S2CONV=1
