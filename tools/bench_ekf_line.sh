#!/bin/bash
python bench.py --workload ${1:-ekf} --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('${1:-ekf} ${2:-}', '%.4g %s  step %.3f ms  e2e %.4g'%(j['value'], j['unit'], j['ms_per_step'], j['e2e']['value']))"
