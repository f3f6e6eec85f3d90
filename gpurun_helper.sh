set -x
timeout 900 python -m pytest tests/test_gpu_next.py tests/test_facade.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --workload safefusion --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r02b_bench_safefusion.json 2> gpurun_out/r02b_bench_safefusion.err
python -c "
import json
d=json.loads(open('gpurun_out/r02b_bench_safefusion.json').read().strip().splitlines()[-1]); print('safefusion', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
