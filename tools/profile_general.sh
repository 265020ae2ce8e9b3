#!/bin/bash
# ncu --set full capture of the CTA-per-window kernel on one workload: tools/profile_general.sh WORKLOAD TAG
# (end-to-end leg in one chunk and without graph replay, so that every solve_kernel launch is the full batch)
w=$1; tag=${2:-gen}
python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --e2e-chunks 1 --no-e2e-graph > gpurun_out/bench_$w.log 2>&1 || { tail -5 gpurun_out/bench_$w.log; exit 1; }
tail -1 gpurun_out/bench_$w.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 3 -c 1 -f -o gpurun_out/prof_${w}_${tag} \
   python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline --e2e-chunks 1 --no-e2e-graph > gpurun_out/ncu_$w.log 2>&1 || tail -3 gpurun_out/ncu_$w.log
