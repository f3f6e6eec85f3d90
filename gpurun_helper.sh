timeout 900 python -m pytest tests/test_gpu_msckf_ekf.py -m gpu -x -q 2>&1 | tail -3
