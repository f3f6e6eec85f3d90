set +x
mkdir -p gpurun_out/scale2
for w in ukfom usckf msckf; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload $w > gpurun_out/scale2/bench_$w.json 2> gpurun_out/scale2/bench_$w.err
python -c "
import json
d=json.loads(open('gpurun_out/scale2/bench_$w.json').read().strip().splitlines()[-1]); print('$w', 'n_gpus', d['n_gpus'], 'value %.4g'%d['value'], 'ms %.4f'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], d['clocks'])"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/scale2/bench_reference.json 2> gpurun_out/scale2/bench_reference.err; cut -c1-200 gpurun_out/scale2/bench_reference.json
