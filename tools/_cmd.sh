timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python tools/configs_bench.py c5 2>&1 | tail -1 | cut -c1-300
python tools/kbench.py default
