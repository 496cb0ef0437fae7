#!/bin/bash
# tools/evidence.sh [nstar] [c4] -- regenerates the ncu evidence behind profiles/r02_* on a B200 (one GPU):
#     gpurun --timeout 900 -- 'bash tools/evidence.sh nstar c4'
# then, back in the build container:
#     python tools/ncu_summary.py r02 gpurun_out/r02_launches.csv gpurun_out/r02_full_raw.csv workload=nstar n=1000000
#     python tools/ncu_summary.py r02c4 gpurun_out/r02c4_launches.csv gpurun_out/r02c4_full_raw.csv workload=c4 n=200000
# Every ncu command runs only after the same command line has exited 0 without ncu; numbers printed under ncu are never
# bench values.  The full library's report embeds 60 MB of cubins, so the `--page raw --csv` export is made on the box and
# only that travels back.
set -u
mkdir -p gpurun_out
what="${*:-nstar c4}"
if [[ " $what " == *" nstar "* ]]; then
  B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras"
  $B > gpurun_out/r02_plain.json 2> gpurun_out/r02_plain.err || { echo "bench failed"; tail -3 gpurun_out/r02_plain.err; exit 1; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_ncu_launches.log 2>&1
  ncu --set full --clock-control none -k regex:'z_propose_kernel|z_kernel|moments_kernel|chi_draw_kernel|stats_kernel_tma|stats_final_kernel' \
      -s 24 -c 12 -f -o /tmp/r02_full $B > gpurun_out/r02_ncu_full.log 2>&1
  ncu -i /tmp/r02_full.ncu-rep --page raw --csv > gpurun_out/r02_full_raw.csv 2> /dev/null
fi
if [[ " $what " == *" c4 "* ]]; then
  # BASELINE configuration 4 (covariate-adjusted, ragged grids): the pair cross-Gram kernel.  The ncu passes run at
  # n = 2e5 (generating the 2e8 ragged observations of n = 1e6 takes 100 s per run); the plain n = 1e6 line is
  # profiles/r02_bench_c4_n1e6.json.
  C4="python bench.py --workload c4 --n 200000 --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
  $C4 > gpurun_out/r02c4_plain.json 2> gpurun_out/r02c4_plain.err || { echo "c4 bench failed"; tail -3 gpurun_out/r02c4_plain.err; exit 1; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02c4_launches.csv $C4 > gpurun_out/r02c4_ncu_launches.log 2>&1
  ncu --set full --clock-control none -k regex:'ragged_stats_kernel|z_kernel|chi_kernel|ssr_kernel' -s 8 -c 6 -f -o /tmp/r02c4_full $C4 > gpurun_out/r02c4_ncu_full.log 2>&1
  ncu -i /tmp/r02c4_full.ncu-rep --page raw --csv > gpurun_out/r02c4_full_raw.csv 2> /dev/null
fi
ls -la gpurun_out/r02*
