for w in vdp gnss; do for c in 8 10 12; do for s in 1 0; do
  echo -n "$w min_ctas=$c const_smem=$s : "
  NLPF_LIB=$PWD/tools/libnlpf_c$c.so NLPF_CONST_SMEM=$s python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('%.0f windows/s  kernel %.3f ms conv %.4f'%(j['value'], j['roofline']['kernel_ms'], j['stats']['converged_fraction']))"
done; done; done
