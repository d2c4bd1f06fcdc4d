# r02 late A/B + validation bundle (ONE call) as it ran for profiles/r02b_*: attention "duo" kernel (DV_ATTN_DUO - measured
# slower and REMOVED afterwards, the variable is ignored now) and reordered MixVPR tail (DV_MIX_TAIL) against the previous
# defaults, then smoke / full GPU suite / headline bench / ncu launch list of bench.py with the winning switches exported.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
python - <<'PY'
import json, os, subprocess, sys
runs = [("base",      dict(DV_ATTN_DUO="0", DV_MIX_TAIL="0")),
        ("tail",      dict(DV_ATTN_DUO="0", DV_MIX_TAIL="1")),
        ("duo1",      dict(DV_ATTN_DUO="1", DV_MIX_TAIL="1")),
        ("duo2",      dict(DV_ATTN_DUO="2", DV_MIX_TAIL="1")),
        ("duo1_3000", dict(DV_ATTN_DUO="1", DV_ATTN_TURN_CLK="3000", DV_MIX_TAIL="1"))]
res = {}
for name, env in runs:
    try:
        r = subprocess.run([sys.executable, "bench.py", "--steps", "8", "--warmup", "3", "--no-latency", "--no-cpu-baseline"],
                           env=dict(os.environ, **env), capture_output=True, text=True, timeout=240)
        d = json.loads(r.stdout.strip().splitlines()[-1])
        res[name] = dict(env=env, value=d["value"], e2e=d["e2e"]["value"], stage=d["stage_ms_per_round"], clocks=d.get("clocks"))
        print(name, round(d["value"]), round(d["e2e"]["value"]), {k: round(v, 3) for k, v in d["stage_ms_per_round"].items()}, flush=True)
    except Exception as ex:
        res[name] = dict(env=env, error=str(ex)[:300], stderr=(r.stderr[-600:] if 'r' in dir() else ""))
        print(name, "FAILED", str(ex)[:200], flush=True)
json.dump(res, open("gpurun_out/r02b_ab.json", "w"), indent=1)
ok = {k: v for k, v in res.items() if "value" in v}
base_lg = min(ok[k]["stage"]["lightglue"] for k in ("base", "base2", "tail") if k in ok)
best_duo, best_lg = ("0", None), base_lg
for k in ("duo1", "duo2", "duo1_3000", "duo1_700"):
    if k in ok and ok[k]["stage"]["lightglue"] < best_lg - 0.02:
        best_lg = ok[k]["stage"]["lightglue"]; best_duo = (ok[k]["env"]["DV_ATTN_DUO"], ok[k]["env"].get("DV_ATTN_TURN_CLK"))
tail = "1" if ("tail" in ok and "base" in ok and ok["tail"]["stage"]["mixvpr"] < min(ok["base"]["stage"]["mixvpr"], ok.get("base2", ok["base"])["stage"]["mixvpr"]) - 0.01) else "0"
with open("gpurun_out/r02b_best.env", "w") as f:
    f.write("export DV_ATTN_DUO=%s\n" % best_duo[0])
    if best_duo[1]: f.write("export DV_ATTN_TURN_CLK=%s\n" % best_duo[1])
    f.write("export DV_MIX_TAIL=%s\n" % tail)
print("BEST", best_duo, "tail", tail, "lg", base_lg, "->", best_lg)
PY
cat gpurun_out/r02b_best.env
. gpurun_out/r02b_best.env
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" 2>&1 | tail -n 3 > gpurun_out/r02b_smoke.log; cat gpurun_out/r02b_smoke.log
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -n 12 > gpurun_out/r02b_pytest_gpu.log; tail -n 3 gpurun_out/r02b_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r02b_bench_euroc_full_1gpu.json 2> gpurun_out/r02b_bench.err; tail -c 300 gpurun_out/r02b_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02b_bench_euroc_full_1gpu.json"))
print("FINAL", round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac"],3), {k:round(v,3) for k,v in d["stage_ms_per_round"].items()}, (d.get("latency_b1") or {}).get("p50_ms"), d["clocks"])
PY
for m in 0 1; do DV_ATTN_DBG=1 DV_ATTN_DUO=$m timeout 200 python bench.py --steps 1 --warmup 3 --no-latency --no-cpu-baseline --no-preflight 2>&1 | grep "attn dbg" | tail -n 4; done | tee gpurun_out/r02b_attn_dbg.log
# ncu launch list of the bench command itself (durations under ncu are serialised / cold: shares, not absolutes)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r02b_bench.csv python bench.py --steps 1 --warmup 3 --no-latency --no-cpu-baseline --no-preflight > gpurun_out/r02b_ncu_bench.log 2>&1; tail -n 2 gpurun_out/r02b_ncu_bench.log | cut -c1-300
python tools/launch_summary.py gpurun_out/launches_r02b_bench.csv > gpurun_out/launches_r02b_bench_summary.txt 2>&1; head -n 50 gpurun_out/launches_r02b_bench_summary.txt
