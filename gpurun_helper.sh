# Round-end validation as run on the GPU box:  gpurun --timeout 3000 -- 'bash gpurun_helper.sh'
# (GPU tests, smoke, the default bench (USCKF fleet + nested configs) and the reference arm, the single-workload benches,
#  the ncu launch list of the bench command and one ncu --set full capture per dominant kernel at the bench's own shape;
#  results under gpurun_out/final/)
set +x
O=gpurun_out/final
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -x -q > $O/gpu_tests.log 2>&1; tail -3 $O/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 300 $O/bench_default.json | cut -c1-200
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
for w in ekf msckf_ekf safefusion deadreckon; do
timeout 900 python bench.py --workload $w > $O/bench_$w.json 2> $O/bench_$w.err; tail -c 200 $O/bench_$w.json | cut -c1-120
done
# launch list of the bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_default.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1
# one --set full capture per dominant kernel, at the bench's own batch size
timeout 900 ncu --set full --clock-control none --import-source on -k regex:usckf_step_kernel -s 6 -c 1 -o $O/ncu_usckf python bench.py --no-also --no-cpu-baseline --no-e2e --steps 5 --warmup 3 > $O/ncu_usckf.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:ukf_kernel -s 6 -c 1 -o $O/ncu_ukfom python bench.py --workload ukfom --no-cpu-baseline --no-e2e --steps 5 --warmup 3 > $O/ncu_ukfom.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:msckf_update_kernel -s 4 -c 1 -o $O/ncu_msckf python bench.py --workload msckf --no-cpu-baseline --no-e2e --steps 5 --warmup 3 > $O/ncu_msckf.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:datamodel_kernel -s 6 -c 1 -o $O/ncu_fusion python bench.py --workload fusion --no-cpu-baseline --no-e2e --steps 5 --warmup 3 > $O/ncu_fusion.log 2>&1
ls -la $O | head -40
