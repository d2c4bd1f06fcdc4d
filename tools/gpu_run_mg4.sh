cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-4}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 5 --warmup 3 --no-latency > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; tail -c 400 gpurun_out/r02_bench_${N}gpu.err; python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_${N}gpu.json'))
print(round(d['value']), round(d['e2e']['value']), d.get('parity_multi'), d.get('lg_partner_counts_e2e_rank0'), d['ms_per_step'])
PY
