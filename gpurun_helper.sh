set +x
mkdir -p gpurun_out/scale8
N=$(nvidia-smi -L | wc -l)
echo "gpus: $N"
for w in ukfom; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --workload $w --no-cpu-baseline > gpurun_out/scale8/bench_$w.json 2> gpurun_out/scale8/bench_$w.err
python -c "
import json
d=json.loads(open('gpurun_out/scale8/bench_$w.json').read().strip().splitlines()[-1]); print('$w', 'n_gpus', d['n_gpus'], 'value %.4g'%d['value'], 'ms %.4f'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], d['clocks'])"
done
