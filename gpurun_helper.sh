set -x
timeout 900 python -m pytest tests/test_gpu_next.py tests/test_facade.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --workload deadreckon --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r02d_bench_deadreckon.json 2> gpurun_out/r02d_bench_deadreckon.err
python -c "
import json
d=json.loads(open('gpurun_out/r02d_bench_deadreckon.json').read().strip().splitlines()[-1]); print('deadreckon', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
