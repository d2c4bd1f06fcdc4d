set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest1.log; tail -5 gpurun_out/pytest1.log
timeout 600 python tools/parity_diag.py > gpurun_out/parity_diag.log 2>&1; tail -3 gpurun_out/parity_diag.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_euroc.json 2> gpurun_out/bench_euroc.err; tail -c 600 gpurun_out/bench_euroc.err; head -c 1500 gpurun_out/bench_euroc.json
for c in kitti_50k sp_lg_512 mix_knn_10k; do timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-latency > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; tail -c 400 gpurun_out/bench_$c.err; head -c 600 gpurun_out/bench_$c.json; done
