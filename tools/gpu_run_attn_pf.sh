# r02 final short A/B: L2 prefetch of the next attention item's Q / K / V tiles (DV_ATTN_PREFETCH) + the LightGlue, pipeline
# and exact-parity tests with the winner + DV_ATTN_DBG counters.  Lands in gpurun_out/r02d_*.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python - <<'PY'
import json, os, subprocess, sys
runs = [("base", dict(DV_ATTN_PREFETCH="0")), ("pf", dict()), ("base_b", dict(DV_ATTN_PREFETCH="0")), ("pf_b", dict())]
res = {}
for name, env in runs:
    try:
        r = subprocess.run([sys.executable, "bench.py", "--steps", "6", "--warmup", "3", "--no-latency", "--no-cpu-baseline"],
                           env=dict(os.environ, **env), capture_output=True, text=True, timeout=120)
        d = json.loads(r.stdout.strip().splitlines()[-1])
        res[name] = dict(env=env, value=d["value"], e2e=d["e2e"]["value"], stage=d["stage_ms_per_round"], clocks=d.get("clocks"))
        print(name, round(d["value"]), round(d["e2e"]["value"]), {k: round(v, 3) for k, v in d["stage_ms_per_round"].items()}, flush=True)
    except Exception as ex:
        res[name] = dict(env=env, error=str(ex)[:300])
        print(name, "FAILED", str(ex)[:200], flush=True)
    json.dump(res, open("gpurun_out/r02d_ab.json", "w"), indent=1)
ok = {k: v for k, v in res.items() if "value" in v}
lg = lambda *ks: min([ok[k]["stage"]["lightglue"] for k in ks if k in ok] or [1e9])
base, pf = lg("base", "base_b"), lg("pf", "pf_b")
open("gpurun_out/r02d_best.env", "w").write("" if pf < base - 0.01 else "export DV_ATTN_PREFETCH=0\n")
print("BEST prefetch", pf < base - 0.01, "base", base, "pf", pf)
PY
. gpurun_out/r02d_best.env
timeout 300 python -m pytest tests/test_lightglue_gpu.py tests/test_pipeline_gpu.py tests/test_parity_exact_gpu.py -m gpu -q 2>&1 | tail -n 4 > gpurun_out/r02d_pytest_lg.log; tail -n 2 gpurun_out/r02d_pytest_lg.log
DV_ATTN_DBG=1 timeout 100 python bench.py --steps 1 --warmup 3 --no-latency --no-cpu-baseline --no-preflight 2>&1 | grep "attn dbg" | head -n 3 > gpurun_out/r02d_attn_dbg.log; tail -n 1 gpurun_out/r02d_attn_dbg.log
