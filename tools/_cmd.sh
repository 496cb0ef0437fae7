timeout 800 python -m pytest tests/test_multi_gpu.py -x -q 2>&1 | tail -5
