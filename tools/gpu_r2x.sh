#!/bin/bash
# round-2 GPU batch X: bucket_kernel of wave w+1 beside count_kernel of wave w, again, now that three count CTAs
# (64 registers) and one bucket CTA fit an SM together; default k = 7 short reads with 8 warps
mkdir -p gpurun_out
O=gpurun_out/r2x
run() { # workload scale opts...
  w=$1; sc=$2; shift 2; flags=""; for kv in "$@"; do flags="$flags --opt $kv"; done
  timeout 300 python bench.py --workload $w --scale $sc --steps 8 --no-e2e --no-cpu --no-cli --no-per-config $flags 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$w', '$sc', '$*', 'Gb/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'rows1', d['rows_sum_to_one'])"
}
{
run reads100k_k10 1.0
for wv in 2 4 8 16; do for c in 1 2; do
run reads100k_k10 1.0 bucket_waves=$wv bucket_wave_ctas=$c
done; done
run reads100k_k10 1.0
run reads150_k7 1.0
run reads150_k6 1.0
run reads10k_k6 0.3
} > $O.sweep.txt 2>&1
cat $O.sweep.txt
