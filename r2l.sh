set +x
mkdir -p gpurun_out/r2l
timeout 900 python -m pytest tests/test_gpu_usckf.py -x -q > gpurun_out/r2l/usckf_tests.log 2>&1; tail -3 gpurun_out/r2l/usckf_tests.log
timeout 600 python bench.py --no-also --no-cpu-baseline --no-e2e --steps 20 > gpurun_out/r2l/bench.json 2> gpurun_out/r2l/bench.err; python -c "
import json;d=json.load(open('gpurun_out/r2l/bench.json'));print(d['ms_per_step'], d['value'])"
timeout 900 ncu --clock-control none -k regex:usckf_step_kernel -s 1 -c 1 --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg,smsp__inst_executed.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum python profiles/run_kernels.py usckf 2>&1 | grep -A12 "Metric Name" | tail -8
