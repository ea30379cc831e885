#!/bin/bash
# usage: tools/sweep_opt.sh <workload> <scale> "k1=v1 k2=v2" "k1=v3" ...   (runs on the GPU box; one bench line per option set)
w=$1; sc=$2; shift 2
for set in "$@"; do
  flags=""; for kv in $set; do flags="$flags --opt $kv"; done
  python bench.py --workload $w --scale $sc --steps 5 --no-e2e --no-cpu $flags 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$set', 'Gb/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3))"
done
