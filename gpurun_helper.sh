set -x
timeout 600 python -m pytest tests/test_gpu_fusion.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --workload fusion --no-cpu-baseline > gpurun_out/r01f_bench_fusion.json 2> gpurun_out/r01f_bench_fusion.err
cut -c1-200 gpurun_out/r01f_bench_fusion.json
