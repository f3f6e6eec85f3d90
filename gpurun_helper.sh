set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r01_final_tests.log
for w in ukfom usckf msckf fusion ekf msckf_ekf safefusion deadreckon; do
timeout 900 python bench.py --workload $w > gpurun_out/r01_final_bench_$w.json 2> gpurun_out/r01_final_bench_$w.err
tail -1 gpurun_out/r01_final_bench_$w.err | cut -c1-200
done
timeout 600 python bench.py --impl reference > gpurun_out/r01_final_ref_ukfom.json 2>/dev/null
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_ukfom.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
wc -l gpurun_out/r01_launches_ukfom.csv
