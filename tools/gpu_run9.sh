cd $GRAFT_REPO_ROOT
DV_ATTN_DBG=1 timeout 300 python bench.py --steps 2 --warmup 3 --no-latency --no-cpu-baseline 2>&1 >/dev/null | grep "attn dbg" | tail -4
