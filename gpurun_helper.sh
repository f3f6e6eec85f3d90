set -x
timeout 900 python -m pytest tests/test_gpu_usckf.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --workload usckf --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r01y_bench_usckf.json 2> gpurun_out/r01y_bench_usckf.err
tail -2 gpurun_out/r01y_bench_usckf.err; cut -c1-220 gpurun_out/r01y_bench_usckf.json
