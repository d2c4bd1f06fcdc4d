# per-CTA timeline of one self- and one cross-attention launch inside the bench pipeline (DV_ATTN_DBG), with and without PDL
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for pdl in 1 0; do
  echo "== DV_PDL=$pdl"
  DV_PDL=$pdl DV_ATTN_DBG=1 timeout 200 python bench.py --steps 1 --warmup 3 --no-latency --no-cpu-baseline --no-preflight 2>&1 | grep "attn" | head -n 24
done | tee gpurun_out/r02b_attn_timeline.log
