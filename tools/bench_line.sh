#!/bin/bash
# one-line bench summary: tools/bench_line.sh WORKLOAD [label]
python bench.py --workload $1 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$1 ${2:-default}', '%.0f %s  step %.3f ms  frac %.3f conv %s'%(j['value'], j['unit'], j['ms_per_step'], j['roofline']['frac'], j.get('stats',{}).get('converged_fraction')))"
