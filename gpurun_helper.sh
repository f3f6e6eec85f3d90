# Round-end validation as run on the GPU box:  gpurun --timeout 1800 -- 'bash gpurun_helper.sh'
# (GPU tests, smoke, every bench workload + the reference arm, the ncu launch list; results under gpurun_out/final/)
set +x
mkdir -p gpurun_out/final
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final/gpu_tests.log 2>&1; tail -3 gpurun_out/final/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final/smoke.log 2>&1; tail -2 gpurun_out/final/smoke.log
for w in ukfom usckf msckf fusion ekf msckf_ekf safefusion deadreckon; do
timeout 900 python bench.py --workload $w > gpurun_out/final/bench_$w.json 2> gpurun_out/final/bench_$w.err; tail -c 300 gpurun_out/final/bench_$w.json | cut -c1-120
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final/bench_reference.json 2>gpurun_out/final/bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final/launches_ukfom.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/final/ncu_launches.log 2>&1
