#!/bin/bash
# round-2 GPU batch L: step loop with alternating prefetch buffers; launch lists of the default step, k = 7 and contigs
mkdir -p gpurun_out
O=gpurun_out/r2l
timeout 1500 python -m pytest tests/test_gpu_long.py tests/test_gpu_parity.py -m gpu -x -q > $O.pytest.txt 2>&1; echo "rc=$?" >> $O.pytest.txt
tail -4 $O.pytest.txt
run() { # workload scale opts...
  w=$1; sc=$2; shift 2; flags=""; for kv in "$@"; do flags="$flags --opt $kv"; done
  timeout 300 python bench.py --workload $w --scale $sc --steps 5 --no-e2e --no-cpu --no-cli --no-per-config $flags 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$w', '$sc', '$*', 'Gb/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'rows1', d['rows_sum_to_one'])"
}
{
run contigs_k4 1.0
run reads10k_k7 1.0
run reads10k_k7 1.0 long_warps=8
run reads10k_k5 0.3
run reads150_k7 1.0
} > $O.sweep.txt 2>&1
cat $O.sweep.txt
KRE='regex:short_kernel|seq_kernel|long_kernel|bucket_kernel|count_kernel|tile_|format_norm|rebase_offsets|wave_kernel|finalize'
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$KRE" -c 12 --csv --log-file $O.launches_k5.csv \
  python bench.py --scale 0.2 --steps 2 --warmup 3 --no-e2e --no-cpu --no-cli --no-per-config > $O.ncu_k5.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$KRE" -c 6 --csv --log-file $O.launches_k7.csv \
  python bench.py --workload reads10k_k7 --scale 0.2 --steps 2 --warmup 3 --no-e2e --no-cpu --no-cli --no-per-config > $O.ncu_k7.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$KRE" -c 8 --csv --log-file $O.launches_contigs.csv \
  python bench.py --workload contigs_k4 --steps 2 --warmup 3 --no-e2e --no-cpu --no-cli --no-per-config > $O.ncu_contigs.log 2>&1
