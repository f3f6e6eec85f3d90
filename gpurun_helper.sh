set -x
timeout 600 python bench.py --workload ekf --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02n_bench_ekf.json 2>gpurun_out/r02n_bench_ekf.err
tail -2 gpurun_out/r02n_bench_ekf.err
python -c "
import json
d=json.loads(open('gpurun_out/r02n_bench_ekf.json').read().strip().splitlines()[-1]); print('ekf', 'value %.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'])"
