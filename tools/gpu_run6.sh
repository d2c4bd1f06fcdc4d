cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
run() { name=$1; shift; env "$@" timeout 300 $B > gpurun_out/sweep_$name.json 2> gpurun_out/sweep_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sweep_$name.json")); print("$name", round(d["value"]), {k:round(v,3) for k,v in d["stage_ms_per_round"].items()}, d.get("p50_match_ms"), d["latency_b1"]["p50_ms"]["lightglue"], d["latency_b1"]["p50_ms"]["total"])
except Exception as ex: print("$name FAILED", ex)
PY
tail -n 3 gpurun_out/sweep_$name.err
}
run ffn0_0 DV_LG_FUSE_FFN0=0
run ffn0_1 DV_LG_FUSE_FFN0=1
