set +x
mkdir -p gpurun_out/final
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final/gpu_tests.log 2>&1; tail -3 gpurun_out/final/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final/smoke.log 2>&1; tail -2 gpurun_out/final/smoke.log
for w in ukfom usckf msckf fusion ekf msckf_ekf safefusion deadreckon; do
timeout 900 python bench.py --workload $w > gpurun_out/final/bench_$w.json 2> gpurun_out/final/bench_$w.err; tail -c 300 gpurun_out/final/bench_$w.json | cut -c1-150
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final/bench_reference.json 2>gpurun_out/final/bench_reference.err
for kk in msckf:msckf_update_kernel msckf_ekf:msckf_ekf_update_kernel; do
arg=${kk%%:*}; kn=${kk#*:}
timeout 400 ncu --set full --clock-control none -k regex:$kn -c 1 -o gpurun_out/final/full_$arg -f python profiles/run_kernels.py $arg > gpurun_out/final/ncu_full_$arg.log 2>&1
python profiles/summarize_ncu.py gpurun_out/final/full_$arg.ncu-rep > gpurun_out/final/ncu_summary_$arg.txt 2>&1
done
