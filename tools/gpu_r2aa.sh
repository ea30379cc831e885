#!/bin/bash
# round-2 GPU batch AA: longest-first work order for long contigs (A/B by option on one box), e2e with the offsets
# validated per chunk (A = previous commit's library)
mkdir -p gpurun_out
O=gpurun_out/r2aa
timeout 900 python -m pytest tests/test_gpu_long.py -m gpu -x -q > $O.pytest.txt 2>&1; echo "rc=$?" >> $O.pytest.txt
tail -3 $O.pytest.txt
run() { # workload scale opts...
  w=$1; sc=$2; shift 2; flags=""; for kv in "$@"; do flags="$flags --opt $kv"; done
  timeout 300 python bench.py --workload $w --scale $sc --steps 8 --no-e2e --no-cpu --no-cli --no-per-config $flags 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$w', '$sc', '$*', 'Gb/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'rows1', d['rows_sum_to_one'])"
}
{
for rep in 1 2; do
run contigs_k4 1.0 longest_first=0
run contigs_k4 1.0 longest_first=1
run contigs_k4 0.25 longest_first=0
run contigs_k4 0.25 longest_first=1
done
} > $O.sweep.txt 2>&1
cat $O.sweep.txt
cp kmertools_b200/lib/libkmertools_b200.so /tmp/libB.so
for L in A B A B; do
  if [ $L = A ]; then cp tools/_probe/libA.so kmertools_b200/lib/libkmertools_b200.so; else cp /tmp/libB.so kmertools_b200/lib/libkmertools_b200.so; fi
  timeout 600 python bench.py --steps 3 --no-cpu --no-cli --no-per-config 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); e=d['e2e']; print('$L e2e', round(e['value'],3), 'ms', round(e['ms_per_step'],1), 'min', round(e['ms_min'],1), 'd2h_ms', round(e['d2h_ms'],1), 'pcie_frac', round(e['pcie_frac'],3))"
done > $O.e2e.txt 2>&1
cat $O.e2e.txt
