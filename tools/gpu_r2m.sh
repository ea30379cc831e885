#!/bin/bash
# round-2 GPU batch M: long_kernel with the two-sequence look-ahead (ticket -> offsets -> bases pipelined across sequences)
mkdir -p gpurun_out
O=gpurun_out/r2m
timeout 1500 python -m pytest tests/test_gpu_long.py tests/test_gpu_parity.py tests/test_gpu_hypothesis.py -m gpu -x -q > $O.pytest.txt 2>&1; echo "rc=$?" >> $O.pytest.txt
tail -4 $O.pytest.txt
run() { # workload scale opts...
  w=$1; sc=$2; shift 2; flags=""; for kv in "$@"; do flags="$flags --opt $kv"; done
  timeout 300 python bench.py --workload $w --scale $sc --steps 5 --no-e2e --no-cpu --no-cli --no-per-config $flags 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$w', '$sc', '$*', 'Gb/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'rows1', d['rows_sum_to_one'])"
}
{
run reads10k_k7 1.0
run reads10k_k7 1.0 long_warps=8
run reads10k_k7 1.0 seq_grab=4
run contigs_k4 1.0
run reads10k_k5 0.3
run reads150_k7 1.0
} > $O.sweep.txt 2>&1
cat $O.sweep.txt
