set -x
timeout 600 python -m pytest tests/test_gpu_msckf_ekf.py tests/test_gpu_msckf.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --workload msckf_ekf --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r01t_bench_msckf_ekf.json 2> gpurun_out/r01t_bench_msckf_ekf.err
tail -3 gpurun_out/r01t_bench_msckf_ekf.err; cut -c1-300 gpurun_out/r01t_bench_msckf_ekf.json
