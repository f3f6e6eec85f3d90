set -x
timeout 900 python -m pytest tests/test_gpu_fusion.py tests/test_facade.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --workload fusion --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r02k_bench_fusion.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/r02k_bench_fusion.json').read().strip().splitlines()[-1]); print('fusion', 'value %.4g'%d['value'], 'ms %.4f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'])"
