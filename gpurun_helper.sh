set -x
timeout 900 python -m pytest tests/test_gpu_usckf.py tests/test_gpu_msckf.py tests/test_facade.py -m gpu -x -q 2>&1 | tail -15
timeout 300 python bench.py --workload usckf --no-cpu-baseline > gpurun_out/r01g_bench_usckf.json 2> gpurun_out/r01g_bench_usckf.err
cut -c1-200 gpurun_out/r01g_bench_usckf.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"usckf_update|predict12" -c 2 -o gpurun_out/prof_r01g_usckf python profiles/run_kernels.py usckf > gpurun_out/ncu_r01g_usckf.log 2>&1
