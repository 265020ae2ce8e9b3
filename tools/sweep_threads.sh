for w in vdp gnss; do for t in 64 128; do for c in 0 1; do
  echo -n "$w threads=$t const_smem=$c : "
  NLPF_THREADS=$t NLPF_CONST_SMEM=$c python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('%.0f windows/s  kernel %.3f ms conv %.4f'%(j['value'], j['roofline']['kernel_ms'], j['stats']['converged_fraction']))"
done; done; done
