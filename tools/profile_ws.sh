#!/bin/bash
# bench (no profiler) then one ncu --set full capture of the warp-per-window kernel per workload
mkdir -p gpurun_out
for w in ${WORKLOADS:-vdp gnss car}; do
  python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.log 2>&1 || { echo "bench $w failed"; tail -5 gpurun_out/bench_$w.log; continue; }
  tail -1 gpurun_out/bench_$w.log | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$w %.0f %s  step %.3f ms  frac %.3f'%(j['value'], j['unit'], j['ms_per_step'], j['roofline']['frac']))"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:wsolve_kernel -s 3 -c 1 -f -o gpurun_out/prof_${w}_${TAG:-ws} \
     python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$w.log 2>&1 || tail -3 gpurun_out/ncu_$w.log
done
