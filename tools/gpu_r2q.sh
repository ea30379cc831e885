#!/bin/bash
# round-2 GPU batch Q: look-ahead for k = 7 only; count_kernel without un-count; then a PROBE build of the library
# (tools/_probe, -DKTB_COUNT_PROBE) with phases of count_kernel switched off, to attribute its time
mkdir -p gpurun_out
O=gpurun_out/r2q
timeout 1500 python -m pytest tests/test_gpu_long.py tests/test_gpu_parity.py tests/test_gpu_bucket.py tests/test_gpu_hypothesis.py -m gpu -x -q > $O.pytest.txt 2>&1; echo "rc=$?" >> $O.pytest.txt
tail -4 $O.pytest.txt
run() { # workload scale opts...
  w=$1; sc=$2; shift 2; flags=""; for kv in "$@"; do flags="$flags --opt $kv"; done
  timeout 300 python bench.py --workload $w --scale $sc --steps 5 --no-e2e --no-cpu --no-cli --no-per-config $flags 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$w', '$sc', '$*', 'probe=$KTB_COUNT_PROBE', 'Gb/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'rows1', d['rows_sum_to_one'])"
}
{
run reads10k_k7 1.0
run contigs_k4 1.0
run reads10k_k5 0.3
run reads150_k7 1.0
run reads100k_k10 1.0
run reads100k_k10_f32 1.0
run reads100k_k9 1.0
} > $O.sweep.txt 2>&1
cat $O.sweep.txt
# ---- probe build (rows are wrong by construction; timing only)
cp tools/_probe/libkmertools_b200.so kmertools_b200/lib/libkmertools_b200.so
{
for pr in 0 1 2 4 8 3 9 11 15 7; do
  export KTB_COUNT_PROBE=$pr
  run reads100k_k10 1.0
done
} > $O.probe.txt 2>&1
cat $O.probe.txt
