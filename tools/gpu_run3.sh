set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest3.log; tail -8 gpurun_out/pytest3.log
timeout 600 python tools/parity_diag.py > gpurun_out/parity_diag.log 2>&1; tail -3 gpurun_out/parity_diag.log
B="python bench.py --steps 5 --warmup 3 --no-latency --no-cpu-baseline"
run() { name=$1; shift; env "$@" timeout 300 $B > gpurun_out/sweep_$name.json 2> gpurun_out/sweep_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sweep_$name.json")); print("$name", round(d["value"]), {k:round(v,3) for k,v in d["stage_ms_per_round"].items()})
except Exception as ex: print("$name FAILED", ex)
PY
}
run base DV_X=0
run nmsL DV_NMS_TILE=L
run mix8 DV_MIX_CHUNK=8
run mix16 DV_MIX_CHUNK=16
run mix32 DV_MIX_CHUNK=32
run lg16 DV_LG_CHUNK=16
run lg32 DV_LG_CHUNK=32
run knn0 DV_KNN_FUSED=0
run nofold DV_LG_FOLD_OUT=0
timeout 300 python bench.py --config mix_knn_10k --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mix_knn_10k.json 2> gpurun_out/bench_mix_knn_10k.err; python -c "
import json; d=json.load(open('gpurun_out/bench_mix_knn_10k.json')); print('mixknn', round(d['value']), d['roofline']['achieved'], d['roofline']['avg_launch_ms'], d['stage_ms_per_round'])"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_euroc.json 2> gpurun_out/bench_euroc.err; python -c "
import json; d=json.load(open('gpurun_out/bench_euroc.json')); print('euroc', round(d['value']), round(d['e2e']['value']), d['p50_match_ms'], json.dumps(d['latency_b1']['p50_ms']))"
