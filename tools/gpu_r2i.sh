#!/bin/bash
# round-2 GPU batch I: wave-overlapped bucket and count kernels
mkdir -p gpurun_out
O=gpurun_out/r2i
timeout 1500 python -m pytest tests/test_gpu_bucket.py tests/test_gpu_long.py -m gpu -x -q > $O.pytest_new.txt 2>&1; echo "rc=$?" >> $O.pytest_new.txt
tail -4 $O.pytest_new.txt
run() { # workload scale opts...
  w=$1; sc=$2; shift 2; flags=""; for kv in "$@"; do flags="$flags --opt $kv"; done
  timeout 300 python bench.py --workload $w --scale $sc --steps 5 --no-e2e --no-cpu --no-cli --no-per-config $flags 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$w', '$sc', '$*', 'Gb/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'rows1', d['rows_sum_to_one'])"
}
{
run reads100k_k10 1.0 bucket_waves=1
run reads100k_k10 1.0 bucket_waves=2
run reads100k_k10 1.0 bucket_waves=4
run reads100k_k10 1.0 bucket_waves=8
run reads100k_k10 1.0 bucket_waves=16
run reads100k_k10_f32 1.0 bucket_waves=8
run reads100k_k9 1.0 bucket_waves=8
run reads100k_k9 1.0 bucket_waves=1
run reads10k_k8 1.0
} > $O.sweep.txt 2>&1
cat $O.sweep.txt
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"bucket_kernel|count_kernel|tile_" -c 4 --csv --log-file $O.launches_k10.csv \
  python bench.py --workload reads100k_k10 --steps 1 --warmup 3 --no-e2e --no-cpu --no-cli --no-per-config > $O.ncu_k10.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O.pytest_all.txt 2>&1; echo "pytest_all rc=$?" >> $O.pytest_all.txt
tail -4 $O.pytest_all.txt
