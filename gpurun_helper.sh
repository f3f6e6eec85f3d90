set -x
timeout 900 python -m pytest tests/test_gpu_msckf.py tests/test_gpu_msckf_ekf.py tests/test_gpu_ukf.py -m gpu -x -q 2>&1 | tail -5
for w in msckf msckf_ekf ukfom; do
timeout 600 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r01v_bench_$w.json 2> gpurun_out/r01v_bench_$w.err
tail -2 gpurun_out/r01v_bench_$w.err; cut -c1-220 gpurun_out/r01v_bench_$w.json
done
