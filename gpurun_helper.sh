set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for w in ukfom usckf msckf fusion; do
timeout 400 python bench.py --workload $w > gpurun_out/r01j_bench_$w.json 2> gpurun_out/r01j_bench_$w.err
cut -c1-250 gpurun_out/r01j_bench_$w.json; tail -2 gpurun_out/r01j_bench_$w.err
done
