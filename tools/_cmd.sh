timeout 600 python tools/configs_bench.py cpo 2>&1 | tail -1
