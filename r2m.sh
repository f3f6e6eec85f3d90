set +x
mkdir -p gpurun_out/r2m
timeout 1200 python -m pytest tests/test_gpu_ukf.py -x -q > gpurun_out/r2m/ukf_tests.log 2>&1; tail -5 gpurun_out/r2m/ukf_tests.log
timeout 600 python bench.py --workload ukfom --no-also --no-cpu-baseline --steps 200 > gpurun_out/r2m/bench_ukfom.json 2> gpurun_out/r2m/bench_ukfom.err; python -c "
import json;d=json.load(open('gpurun_out/r2m/bench_ukfom.json'));print(d['ms_per_step'], d['value'], d['e2e']['value'])"
