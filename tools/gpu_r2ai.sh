#!/bin/bash
# round-2 GPU batch AI: k = 6 inside long_kernel (MODE_CAN); A = previous commit
mkdir -p gpurun_out
O=gpurun_out/r2ai
cp kmertools_b200/lib/libkmertools_b200.so /tmp/libB.so
timeout 900 python -m pytest tests/test_gpu_long.py tests/test_gpu_parity.py tests/test_gpu_hypothesis.py -m gpu -x -q > $O.pytest.txt 2>&1; echo "rc=$?" >> $O.pytest.txt
tail -5 $O.pytest.txt
run() { # tag workload scale opts...
  tag=$1; w=$2; sc=$3; shift 3; flags=""; for kv in "$@"; do flags="$flags --opt $kv"; done
  timeout 300 python bench.py --workload $w --scale $sc --steps 8 --no-e2e --no-cpu --no-cli --no-per-config $flags 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$tag', '$w', '$sc', '$*', 'Gb/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'rows1', d['rows_sum_to_one'])"
}
{
cp tools/_probe/libA.so kmertools_b200/lib/libkmertools_b200.so
run A reads150_k6 1.0
run A reads10k_k6 0.3
cp /tmp/libB.so kmertools_b200/lib/libkmertools_b200.so
run B reads150_k6 1.0
run B reads150_k6 1.0 long_warps=8
run B reads10k_k6 0.3
run B reads10k_k6 0.3 long_warps=8
run B reads150_k6 1.0 k6_long=0
} > $O.sweep.txt 2>&1
cat $O.sweep.txt
