set -x
timeout 900 python -m pytest tests/test_gpu_next.py tests/test_facade.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --workload ekf --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r02i_bench_ekf.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/r02i_bench_ekf.json').read().strip().splitlines()[-1]); print('ekf', 'value %.4g'%d['value'], 'ms %.3f'%d['ms_per_step'], 'upd %.3f'%d['roofline']['kernel_ms'], 'frac %.3f'%d['roofline']['frac'])"
