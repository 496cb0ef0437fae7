timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -x -q 2>&1 | tail -3
python tools/kbench.py default
