set +x
timeout 900 python -m pytest tests/test_gpu_msckf.py tests/test_gpu_msckf_ekf.py -m gpu -x -q 2>&1 | tail -3
for w in msckf msckf_ekf; do
timeout 600 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/tmp_bench.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/tmp_bench.json').read().strip().splitlines()[-1]); print('$w', 'value %.4g'%d['value'], 'ms %.3f'%d['ms_per_step'])"
done
cp slam-localization_b200/csrc/libslb.so /tmp/libslb_orig.so
for cw in 0:ukf 1:ukf; do
c=${cw%%:*}; w=${cw#*:}
make -C slam-localization_b200/csrc timing CALL=$c > /dev/null 2>&1
cp slam-localization_b200/csrc/libslb_timing.so slam-localization_b200/csrc/libslb.so
echo "== chol_blocked call $c of the $w flavour"
timeout 120 python profiles/chol_timing.py $w slam-localization_b200/csrc/libslb.so 2>&1 | tail -15
cp /tmp/libslb_orig.so slam-localization_b200/csrc/libslb.so
done
