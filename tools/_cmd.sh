timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
export BFMMM_LIB=$PWD/tools/lib_small.so
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none -k regex:'z_kernel|chi_kernel|ssr_kernel|stats_kernel' -s 20 -c 4 -o gpurun_out/prof_r1_final python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f2.log 2>&1
tail -1 gpurun_out/ncu_f2.log | cut -c1-80
