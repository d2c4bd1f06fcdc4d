cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-latency --no-cpu-baseline"
run() { name=$1; shift; env "$@" timeout 300 $B > gpurun_out/sweep_$name.json 2> gpurun_out/sweep_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sweep_$name.json")); print("$name", round(d["value"]), {k:round(v,3) for k,v in d["stage_ms_per_round"].items()}, round(d["roofline"]["frac"],3))
except Exception as ex: print("$name FAILED", ex)
PY
tail -n 2 gpurun_out/sweep_$name.err
}
run pref0 DV_GEMM_PAIR_PREFETCH=0
run pref1 DV_GEMM_PAIR_PREFETCH=1
