cd $GRAFT_REPO_ROOT
DV_FFN0_DBG=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-latency --no-cpu-baseline --no-preflight 2>&1 >/dev/null | grep "ffn0 dbg" | tail -3
