#!/bin/bash
# round-end evidence: bench lines (no profiler), then the ncu launch list of the default bench command and
# one ncu --set full capture per warp-kernel workload
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err || tail -5 gpurun_out/bench_default.err
tail -1 gpurun_out/bench_default.json | cut -c1-300
for w in gnss car 2rx ekf ls; do python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err || tail -3 gpurun_out/bench_$w.err; tail -1 gpurun_out/bench_$w.json | cut -c1-200; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_vdp.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1 || tail -3 gpurun_out/ncu_launches.log
TAG=final tools/profile_ws.sh > gpurun_out/profile_ws_final.log 2>&1
tail -3 gpurun_out/profile_ws_final.log
