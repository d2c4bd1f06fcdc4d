cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-latency --no-cpu-baseline"
run() { name=$1; shift; env "$@" timeout 300 $B > gpurun_out/sweep_$name.json 2> gpurun_out/sweep_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sweep_$name.json")); print("$name", round(d["value"]), {k:round(v,3) for k,v in d["stage_ms_per_round"].items()})
except Exception as ex: print("$name FAILED", ex)
PY
}
run pair0 DV_GEMM_PAIR=0
run pair2n0 DV_GEMM_PAIR=2 DV_GEMM_PAIR_NARROW=0
run pair2n1 DV_GEMM_PAIR=2 DV_GEMM_PAIR_NARROW=1
run pair2n2 DV_GEMM_PAIR=2 DV_GEMM_PAIR_NARROW=2
