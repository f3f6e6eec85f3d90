set -x
timeout 900 python -m pytest tests/test_gpu_msckf_ekf.py tests/test_gpu_next.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --workload msckf_ekf --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02q_bench_msckf_ekf.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/r02q_bench_msckf_ekf.json').read().strip().splitlines()[-1]); print('msckf_ekf', 'value %.4g'%d['value'], 'ms %.3f'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'])"
