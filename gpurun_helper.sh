set -x
timeout 900 python -m pytest tests/test_gpu_usckf.py -m gpu -x -q 2>&1 | tail -8
timeout 300 python bench.py --workload usckf --no-cpu-baseline --no-e2e --steps 30 --warmup 5 > gpurun_out/r01p_bench_usckf_m3.json 2> gpurun_out/r01p_bench_usckf_m3.err
SLB_USCKF_MINB=4 timeout 300 python bench.py --workload usckf --no-cpu-baseline --no-e2e --steps 30 --warmup 5 > gpurun_out/r01p_bench_usckf_m4.json 2> gpurun_out/r01p_bench_usckf_m4.err
cut -c1-400 gpurun_out/r01p_bench_usckf_m3.json; cut -c1-400 gpurun_out/r01p_bench_usckf_m4.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"predict12|usckf_update" -c 2 -o gpurun_out/prof_r01p_usckf python profiles/run_kernels.py usckf > gpurun_out/ncu_r01p.log 2>&1
tail -2 gpurun_out/ncu_r01p.log
