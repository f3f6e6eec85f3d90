set +x
mkdir -p gpurun_out/final
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final/gpu_tests.log 2>&1; tail -3 gpurun_out/final/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final/smoke.log 2>&1; tail -2 gpurun_out/final/smoke.log
for w in ukfom msckf msckf_ekf; do
timeout 900 python bench.py --workload $w > gpurun_out/final/bench_$w.json 2> gpurun_out/final/bench_$w.err; tail -c 300 gpurun_out/final/bench_$w.json | cut -c1-120
done
for kk in msckf:msckf_update_kernel msckf_ekf:msckf_ekf_update_kernel; do
arg=${kk%%:*}; kn=${kk#*:}
timeout 400 ncu --set full --clock-control none -k regex:$kn -c 1 -o gpurun_out/final/full_$arg -f python profiles/run_kernels.py $arg > gpurun_out/final/ncu_full_$arg.log 2>&1
python profiles/summarize_ncu.py gpurun_out/final/full_$arg.ncu-rep > gpurun_out/final/ncu_summary_$arg.txt 2>&1
rm -f gpurun_out/final/full_$arg.ncu-rep
done
cp slam-localization_b200/csrc/libslb.so /tmp/libslb_orig.so
for cw in 0:ukf 1:ukf 0:ekf 1:ekf; do
c=${cw%%:*}; w=${cw#*:}
make -C slam-localization_b200/csrc timing CALL=$c > /dev/null 2>&1
cp slam-localization_b200/csrc/libslb_timing.so slam-localization_b200/csrc/libslb.so
echo "== chol_blocked call $c of the $w flavour" >> gpurun_out/final/chol_timeline.txt
timeout 120 python profiles/chol_timing.py $w slam-localization_b200/csrc/libslb.so >> gpurun_out/final/chol_timeline.txt 2>&1
cp /tmp/libslb_orig.so slam-localization_b200/csrc/libslb.so
done
grep total gpurun_out/final/chol_timeline.txt
