set +x
timeout 900 python -m pytest tests/test_gpu_msckf.py tests/test_gpu_msckf_ekf.py -m gpu -x -q 2>&1 | tail -3
for w in msckf msckf_ekf; do
timeout 600 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/tmp_bench.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/tmp_bench.json').read().strip().splitlines()[-1]); print('$w', 'value %.4g'%d['value'], 'ms %.3f'%d['ms_per_step'])"
done
