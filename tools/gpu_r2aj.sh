#!/bin/bash
# round-2 GPU batch AJ: count_kernel with 16 sequences per unit (chunk-major order); B = current library, C = -DKTB_CK_SEQ_CHUNK=16
mkdir -p gpurun_out
O=gpurun_out/r2aj
cp kmertools_b200/lib/libkmertools_b200.so /tmp/libB.so
run() { # tag workload scale opts...
  tag=$1; w=$2; sc=$3; shift 3; flags=""; for kv in "$@"; do flags="$flags --opt $kv"; done
  timeout 300 python bench.py --workload $w --scale $sc --steps 8 --no-e2e --no-cpu --no-cli --no-per-config $flags 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$tag', '$w', '$sc', '$*', 'Gb/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'rows1', d['rows_sum_to_one'])"
}
{
for rep in 1 2; do
cp /tmp/libB.so kmertools_b200/lib/libkmertools_b200.so
run B reads100k_k10 1.0
run B reads100k_k10_f32 1.0
cp tools/_probe/libC.so kmertools_b200/lib/libkmertools_b200.so
run C reads100k_k10 1.0
run C reads100k_k10_f32 1.0
done
timeout 600 python -m pytest tests/test_gpu_bucket.py -m gpu -x -q 2>&1 | tail -2
} > $O.sweep.txt 2>&1
cat $O.sweep.txt
