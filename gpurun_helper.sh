for n in 1 2 3 4 6; do
SLB_HOST_CHUNKS=$n timeout 600 python bench.py --workload ukfom --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/tmp_bench.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/tmp_bench.json').read().strip().splitlines()[-1]); print('chunks $n', 'e2e %.4g'%d['e2e']['value'], 'us/step %.1f'%(65536/d['e2e']['value']*1e6))"
done
for n in 2 4 8; do
SLB_HOST_CHUNKS=$n timeout 600 python bench.py --workload usckf --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/tmp_bench.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/tmp_bench.json').read().strip().splitlines()[-1]); print('usckf chunks $n', 'e2e %.4g'%d['e2e']['value'])"
done
