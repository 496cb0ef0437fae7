timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 100 --warmup 10 > gpurun_out/bench_new.log 2>&1; tail -1 gpurun_out/bench_new.log
