set -x
for k in ukf_kernel predict12_kernel usckf_update_kernel datamodel_kernel msckf_update_kernel ekf_predict_kernel ekf_update_kernel msckf_ekf_update_kernel safe_fusion_kernel dr_update_pose_kernel; do
case $k in
 ukf_kernel) w=ukf;; predict12_kernel|usckf_update_kernel) w=usckf;; datamodel_kernel) w=fusion;; msckf_update_kernel) w=msckf;;
 ekf_predict_kernel|ekf_update_kernel) w=ekf;; msckf_ekf_update_kernel) w=msckf_ekf;; *) w=next;;
esac
timeout 300 ncu --set full --clock-control none -k regex:"^$k" -s 1 -c 1 -o gpurun_out/fin_$k python profiles/run_kernels.py $w > gpurun_out/ncu_fin_$k.log 2>&1
tail -1 gpurun_out/ncu_fin_$k.log
done
ls -la gpurun_out/*.ncu-rep | awk '{s+=$5} END {print s/1e6 " MB"}'
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_ukfom.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 600 python bench.py --workload usckf --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain_bench2.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r01_launches_usckf.csv python bench.py --workload usckf --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches2.log 2>&1
wc -l gpurun_out/r01_launches_*.csv
