timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py > gpurun_out/bench_final.log 2>&1; tail -1 gpurun_out/bench_final.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], {k: round(v['ms']*1e3,1) for k,v in d['roofline']['kernels'].items()}, d['roofline']['frac'], d['roofline']['traffic'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'])"
python bench.py --impl reference --steps 1 --warmup 0 | tail -1 | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
