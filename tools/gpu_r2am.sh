#!/bin/bash
# round-2 GPU batch AM: MODE_K8 with 6 warps per CTA (18 warps per SM)
mkdir -p gpurun_out
O=gpurun_out/r2am
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "k8" > $O.pytest.txt 2>&1; echo "rc=$?" >> $O.pytest.txt
tail -3 $O.pytest.txt
run() { # workload scale opts...
  w=$1; sc=$2; shift 2; flags=""; for kv in "$@"; do flags="$flags --opt $kv"; done
  timeout 300 python bench.py --workload $w --scale $sc --steps 8 --no-e2e --no-cpu --no-cli --no-per-config $flags 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$w', '$sc', '$*', 'Gb/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'rows1', d['rows_sum_to_one'])"
}
{
for rep in 1 2; do
run reads10k_k8 1.0 long_warps=4
run reads10k_k8 1.0 long_warps=6
run reads10k_k8 1.0 long_warps=8
done
} > $O.sweep.txt 2>&1
cat $O.sweep.txt
