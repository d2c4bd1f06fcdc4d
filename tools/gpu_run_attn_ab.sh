# r02 last A/B + validation bundle (ONE call): L2 prefetch of the next attention item's Q tile (DV_ATTN_QPREFETCH) and the
# lazy-rescale threshold (DV_ATTN_GROW) against the previous behaviour, DV_ATTN_DBG counters, then smoke / full GPU
# suite / headline bench with the winning switches exported.  Everything lands in gpurun_out/r02c_*.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python - <<'PY'
import json, os, subprocess, sys
runs = [("base",   dict(DV_ATTN_QPREFETCH="0")),
        ("pf",     dict()),
        ("pf_g12", dict(DV_ATTN_GROW="12")),
        ("base_b", dict(DV_ATTN_QPREFETCH="0")),
        ("pf_b",   dict())]
res = {}
for name, env in runs:
    try:
        r = subprocess.run([sys.executable, "bench.py", "--steps", "8", "--warmup", "3", "--no-latency", "--no-cpu-baseline"],
                           env=dict(os.environ, **env), capture_output=True, text=True, timeout=240)
        d = json.loads(r.stdout.strip().splitlines()[-1])
        res[name] = dict(env=env, value=d["value"], e2e=d["e2e"]["value"], stage=d["stage_ms_per_round"], clocks=d.get("clocks"))
        print(name, round(d["value"]), round(d["e2e"]["value"]), {k: round(v, 3) for k, v in d["stage_ms_per_round"].items()}, flush=True)
    except Exception as ex:
        res[name] = dict(env=env, error=str(ex)[:300])
        print(name, "FAILED", str(ex)[:200], flush=True)
json.dump(res, open("gpurun_out/r02c_ab.json", "w"), indent=1)
ok = {k: v for k, v in res.items() if "value" in v}
lg = lambda *ks: min(ok[k]["stage"]["lightglue"] for k in ks if k in ok)
base, pf = lg("base", "base_b"), lg("pf", "pf_b")
env = {}
if not pf < base - 0.01: env["DV_ATTN_QPREFETCH"] = "0"
best = min(base, pf)
if "pf_g12" in ok and ok["pf_g12"]["stage"]["lightglue"] < best - 0.05: env["DV_ATTN_GROW"] = "12"
open("gpurun_out/r02c_best.env", "w").write("".join("export %s=%s\n" % kv for kv in env.items()))
print("BEST env", env, "base", base, "pf", pf)
PY
cat gpurun_out/r02c_best.env
for m in 0 1; do echo "== DV_ATTN_QPREFETCH=$m"; DV_ATTN_DBG=1 DV_ATTN_QPREFETCH=$m timeout 200 python bench.py --steps 1 --warmup 3 --no-latency --no-cpu-baseline --no-preflight 2>&1 | grep "attn" | head -n 18 | tail -n 17; done > gpurun_out/r02c_attn_dbg.log; grep -h "attn dbg" gpurun_out/r02c_attn_dbg.log | sed -n '2p;12p'
. gpurun_out/r02c_best.env
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" 2>&1 | tail -n 3 > gpurun_out/r02c_smoke.log; cat gpurun_out/r02c_smoke.log
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -n 12 > gpurun_out/r02c_pytest_gpu.log; tail -n 3 gpurun_out/r02c_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r02c_bench_euroc_full_1gpu.json 2> gpurun_out/r02c_bench.err; tail -c 300 gpurun_out/r02c_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02c_bench_euroc_full_1gpu.json"))
print("FINAL", round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac"],3), {k:round(v,3) for k,v in d["stage_ms_per_round"].items()}, (d.get("latency_b1") or {}).get("p50_ms"), d["clocks"])
PY
