timeout 900 python -m pytest tests -m gpu -x -q -k "ragged or cov or Ragged" 2>&1 | tail -3
timeout 900 python tools/configs_bench.py c4 2>&1 | tail -1 | cut -c1-250
