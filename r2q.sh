set +x
mkdir -p gpurun_out/r2q
nvidia-smi topo -m > gpurun_out/r2q/topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2q/bench_8gpu.json 2> gpurun_out/r2q/bench_8gpu.err; tail -c 800 gpurun_out/r2q/bench_8gpu.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2q/bench_8gpu.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['pcie_gbs_per_rank'], 'full', d['e2e_full_posterior']['value'], 'gather', d['ensemble_stats_allreduce_ms'], d['ensemble_stats_instances'])
for k,v in d['also'].items(): print(k, v['value'], v['e2e']['value'], v['ensemble_stats_allreduce_ms'], v['ensemble_stats_instances'])
PY
