cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q -s 2>&1 | tail -n 25 > gpurun_out/r02_pytest_multigpu_2gpu.log; tail -n 6 gpurun_out/r02_pytest_multigpu_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; tail -c 600 gpurun_out/r02_bench_2gpu.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_2gpu.json'))
print(round(d['value']), round(d['e2e']['value']), d.get('parity_multi'), d.get('lg_partner_counts_rank0'), d['ms_per_step'])
PY
