#!/bin/bash
# round-2 evidence: one ncu --set full capture per warp-kernel workload (after the same command ran without the profiler)
mkdir -p gpurun_out
run() {  # name regex args...
  name=$1; rx=$2; shift 2
  python bench.py "$@" --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_pre_$name.log 2>&1 || { echo "bench $name failed"; tail -3 gpurun_out/prof_pre_$name.log; return; }
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -f -o gpurun_out/prof_r2_$name \
     python bench.py "$@" --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r2_$name.log 2>&1 || tail -3 gpurun_out/ncu_r2_$name.log
}
run vdp wsolve_kernel --workload vdp
run gnss wsolve_kernel --workload gnss
run car wsolve_kernel --workload car --chain 25
run asm_vdp ws_assemble_kernel --workload assemble-vdp
run asm_gnss ws_assemble_kernel --workload assemble-gnss
ls -la gpurun_out/prof_r2_*.ncu-rep
