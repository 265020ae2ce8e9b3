#!/bin/bash
# end of round 2 (new damping schedule): launch list of the default command, then one ncu --set full capture of the
# whole-solve kernel for the van der Pol, GNSS and vehicle workloads (each after the same command ran without the profiler)
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_pre_default.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_launches_vdp.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_ncu_launch.log 2>&1
run() {  # name regex args...
  name=$1; rx=$2; shift 2
  python bench.py "$@" --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_pre_$name.log 2>&1 || { echo "bench $name failed"; tail -3 gpurun_out/prof_pre_$name.log; return; }
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -f -o gpurun_out/prof_r2c_$name \
     python bench.py "$@" --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r2c_$name.log 2>&1 || tail -3 gpurun_out/ncu_r2c_$name.log
}
# (one chunk, no graph: every wsolve_kernel launch of the process is then the full 4096-window batch, so -s 3 -c 1 lands on it)
run vdp wsolve_kernel --workload vdp --e2e-chunks 1 --no-e2e-graph
run gnss wsolve_kernel --workload gnss --e2e-chunks 1 --no-e2e-graph
run car wsolve_kernel --workload car --chain 25
ls -la gpurun_out/prof_r2c_*.ncu-rep
