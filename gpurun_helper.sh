set -x
timeout 900 python -m pytest tests/test_gpu_next.py -m gpu -x -q 2>&1 | tail -5
for w in ekf safefusion deadreckon; do
timeout 400 python bench.py --workload $w --steps 50 --warmup 5 > gpurun_out/r01q_bench_$w.json 2> gpurun_out/r01q_bench_$w.err
tail -3 gpurun_out/r01q_bench_$w.err; cut -c1-200 gpurun_out/r01q_bench_$w.json
done
