#!/bin/bash
# ncu --set full capture of one kernel (regex) on one bench workload: tools/profile_one.sh WORKLOAD KERNEL_REGEX TAG
w=$1; k=$2; tag=${3:-one}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_${w}_${tag} \
   python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$w.log 2>&1 || tail -3 gpurun_out/ncu_$w.log
