#!/bin/bash
# every bench line of DESIGN section 5.3, one JSON line each, into gpurun_out/$1
out=gpurun_out/${1:-bench_all.log}
: > $out
for w in vdp gnss car 2rx rc mrx; do timeout 600 python bench.py --workload $w | tail -1 >> $out 2>&1; done
timeout 300 python bench.py --workload car --chain 25 --no-cpu-baseline | tail -1 >> $out 2>&1
