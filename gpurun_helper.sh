set -x
timeout 900 python -m pytest tests/test_gpu_usckf.py tests/test_gpu_msckf.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --workload usckf --no-cpu-baseline --no-e2e > gpurun_out/r01i_bench_usckf.json 2> gpurun_out/r01i_bench_usckf.err
cut -c1-200 gpurun_out/r01i_bench_usckf.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"usckf_update|predict12" -c 4 --csv --log-file gpurun_out/r01i_launches.csv python profiles/run_kernels.py usckf > /dev/null 2>&1
tail -4 gpurun_out/r01i_launches.csv | cut -c1-300
