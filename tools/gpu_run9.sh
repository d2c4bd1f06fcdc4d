cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_lightglue_gpu.py -m gpu -x -q -s -k fused_ffn0 2>&1 | tail -n 12
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_mixvpr_gpu.py tests/test_lightglue_gpu.py tests/test_parity_exact_gpu.py -m gpu -q 2>&1 | tail -n 8
B="python bench.py --steps 5 --warmup 3 --no-latency --no-cpu-baseline"
run() { name=$1; shift; env "$@" timeout 300 $B > gpurun_out/sweep_$name.json 2> gpurun_out/sweep_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sweep_$name.json")); print("$name", round(d["value"]), {k:round(v,3) for k,v in d["stage_ms_per_round"].items()})
except Exception as ex: print("$name FAILED", ex)
PY
tail -n 3 gpurun_out/sweep_$name.err
}
run epi16_0 DV_GEMM_EPI16=0
run epi16_1 DV_GEMM_EPI16=1
