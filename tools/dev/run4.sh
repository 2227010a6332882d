timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -40
