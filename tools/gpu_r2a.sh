#!/bin/bash
# round-2 GPU batch A: parity of long_kernel, DSMEM atomics microbench, old-vs-new sweeps, one ncu capture
mkdir -p gpurun_out
O=gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O.gpu.txt
timeout 900 python -m pytest tests/test_gpu_long.py -m gpu -x -q > $O.pytest_long.txt 2>&1; echo "pytest_long rc=$?" >> $O.pytest_long.txt
tail -3 $O.pytest_long.txt
timeout 120 tools/_build/microbench_dsmem > $O.dsmem.txt 2>&1; tail -12 $O.dsmem.txt
run() { # workload scale opts...
  w=$1; sc=$2; shift 2; flags=""; for kv in "$@"; do flags="$flags --opt $kv"; done
  timeout 300 python bench.py --workload $w --scale $sc --steps 5 --no-e2e --no-cpu --no-cli $flags 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$w', '$sc', '$*', 'Gb/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'rows1', d['rows_sum_to_one'])"
}
{
run reads10k_k7 0.3 k7_mid=0
run reads10k_k7 0.3 k7_mid=1
run reads10k_k7 1.0 k7_mid=1
run reads150_k7 1.0 k7_mid=0
run reads150_k7 1.0 k7_mid=1
run contigs_k4 1.0 fwd_fold=0
run contigs_k4 1.0 fwd_fold=1
run reads10k_k5 0.3 fwd_fold=0
run reads10k_k5 0.3 fwd_fold=1
run reads10k_k6 0.3 fwd_fold=0
run reads10k_k6 0.3 fwd_fold=1
} > $O.sweep.txt 2>&1
cat $O.sweep.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:long_kernel -s 3 -c 1 -o $O.prof_k7 \
  python bench.py --workload reads10k_k7 --scale 0.05 --steps 1 --warmup 3 --no-e2e --no-cpu --no-cli > $O.ncu_k7.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:long_kernel -s 3 -c 1 -o $O.prof_contigs \
  python bench.py --workload contigs_k4 --scale 0.25 --steps 1 --warmup 3 --no-e2e --no-cpu --no-cli > $O.ncu_contigs.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $O.pytest_all.txt 2>&1; echo "pytest_all rc=$?" >> $O.pytest_all.txt
tail -3 $O.pytest_all.txt
