set -x
timeout 900 python -m pytest tests/test_gpu_msckf.py tests/test_facade.py -m gpu -x -q 2>&1 | tail -15
timeout 300 python bench.py --workload msckf --no-cpu-baseline > gpurun_out/r01n_bench_msckf.json 2> gpurun_out/r01n_bench_msckf.err
cut -c1-200 gpurun_out/r01n_bench_msckf.json
