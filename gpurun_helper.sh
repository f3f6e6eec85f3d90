timeout 900 python -m pytest tests/test_gpu_fusion.py -m gpu -x -q 2>&1 | tail -3
