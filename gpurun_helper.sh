set +x
mkdir -p gpurun_out/final
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final/gpu_tests.log 2>&1; tail -3 gpurun_out/final/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final/smoke.log 2>&1; tail -2 gpurun_out/final/smoke.log
for w in ukfom usckf msckf fusion ekf msckf_ekf safefusion deadreckon; do
timeout 900 python bench.py --workload $w > gpurun_out/final/bench_$w.json 2> gpurun_out/final/bench_$w.err; tail -c 400 gpurun_out/final/bench_$w.json | cut -c1-200
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final/bench_reference.json 2>gpurun_out/final/bench_reference.err; cut -c1-300 gpurun_out/final/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final/launches_ukfom.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/final/ncu_launches.log 2>&1
tail -3 gpurun_out/final/launches_ukfom.csv | cut -c1-300
