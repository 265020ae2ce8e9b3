for c in 5 6 7; do
  echo -n "car min_ctas=$c : "
  NLPF_LIB=$PWD/tools/libnlpf_c$c.so python bench.py --workload car --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('%.0f windows/s  step %.3f ms'%(j['value'], j['ms_per_step']))"
  echo -n "gnss-general min_ctas=$c : "
  NLPF_NO_WARP_KERNEL=1 NLPF_LIB=$PWD/tools/libnlpf_c$c.so python bench.py --workload gnss --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('%.0f windows/s  kernel %.3f ms'%(j['value'], j['roofline']['kernel_ms']))"
done
