set +x
mkdir -p gpurun_out/r2b
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b/gpu_tests.log 2>&1; tail -5 gpurun_out/r2b/gpu_tests.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:usckf_step_kernel -s 1 -c 1 -o gpurun_out/r2b/usckf_step python profiles/run_kernels.py usckf > gpurun_out/r2b/ncu.log 2>&1; tail -2 gpurun_out/r2b/ncu.log
