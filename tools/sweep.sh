#!/bin/bash
# usage: tools/sweep.sh <workload> <scale> <option-name> v1 v2 ...   (runs on the GPU box)
w=$1; sc=$2; opt=$3; shift 3
for v in "$@"; do
  python bench.py --workload $w --scale $sc --steps 5 --no-e2e --no-cpu --$opt $v 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$opt', '$v', 'Gb/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3))"
done
