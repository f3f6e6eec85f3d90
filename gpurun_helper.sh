set -x
timeout 900 python -m pytest tests/test_gpu_usckf.py tests/test_gpu_msckf.py tests/test_facade.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --workload usckf --no-cpu-baseline --no-e2e > gpurun_out/r01k_bench_usckf.json 2> gpurun_out/r01k_bench_usckf.err
cut -c1-200 gpurun_out/r01k_bench_usckf.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"usckf_update|predict12" -c 4 --csv --log-file gpurun_out/r01k_launches.csv python profiles/run_kernels.py usckf > /dev/null 2>&1
tail -2 gpurun_out/r01k_launches.csv | cut -c60-300
