set -x
timeout 900 python -m pytest tests/test_gpu_usckf.py tests/test_gpu_ukf.py -m gpu -x -q 2>&1 | tail -4
for z in 1 0; do
SLB_ZERO_COPY=$z timeout 600 python bench.py --workload usckf --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02g_bench_usckf_zc$z.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/r02g_bench_usckf_zc$z.json').read().strip().splitlines()[-1]); print('usckf zero_copy $z', 'value %.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'])"
done
