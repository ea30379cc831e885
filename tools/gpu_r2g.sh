#!/bin/bash
# round-2 GPU batch G: long_kernel without priming (look-back lane), count_kernel pipeline, writer modes, hypothesis parity
mkdir -p gpurun_out
O=gpurun_out/r2g
timeout 1500 python -m pytest tests/test_gpu_bucket.py tests/test_gpu_long.py tests/test_gpu_hypothesis.py -m gpu -x -q > $O.pytest_new.txt 2>&1; echo "rc=$?" >> $O.pytest_new.txt
tail -8 $O.pytest_new.txt
run() { # workload scale opts...
  w=$1; sc=$2; shift 2; flags=""; for kv in "$@"; do flags="$flags --opt $kv"; done
  timeout 300 python bench.py --workload $w --scale $sc --steps 5 --no-e2e --no-cpu --no-cli --no-per-config $flags 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$w', '$sc', '$*', 'Gb/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'rows1', d['rows_sum_to_one'])"
}
{
run reads100k_k10 1.0 bucket=1
run reads100k_k10_f32 1.0 bucket=1
run reads100k_k9 1.0 bucket=1
run reads10k_k7 0.3 k7_mid=0
run reads10k_k7 0.3 k7_mid=1 long_warps=8
run reads10k_k7 0.3 k7_mid=1 long_warps=10
run reads10k_k7 0.3 k7_mid=1 long_warps=4
run reads10k_k7 1.0 k7_mid=1 long_warps=8
run contigs_k4 1.0 fwd_fold=1
run reads10k_k5 0.3 fwd_fold=1 long_warps=4
run reads10k_k5 0.3 fwd_fold=1 long_warps=8
run reads10k_k5 0.3 fwd_fold=0
} > $O.sweep.txt 2>&1
cat $O.sweep.txt
for mode in seq map; do
  KTB_WRITER=$mode KTB_FILE_TRACE=1 python bench.py --steps 2 --no-e2e --no-cpu --no-per-config 2> $O.cli_$mode.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); c=d['cli']; print('$mode', {k:(round(v,3) if isinstance(v,float) else v) for k,v in c.items() if k not in ('what','cpu_baseline')})" >> $O.cli.txt
  grep "ktb file" $O.cli_$mode.err | tail -2 >> $O.cli.txt
done
df -T /tmp | tail -1 >> $O.cli.txt
cat $O.cli.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:long_kernel -s 3 -c 1 -o $O.prof_k7 \
  python bench.py --workload reads10k_k7 --scale 0.05 --steps 1 --warmup 3 --no-e2e --no-cpu --no-cli --no-per-config --opt k7_mid=1 --opt long_warps=8 > $O.ncu_k7.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:count_kernel -s 1 -c 1 -o $O.prof_count \
  python bench.py --workload reads100k_k10 --scale 0.5 --steps 1 --warmup 3 --no-e2e --no-cpu --no-cli --no-per-config > $O.ncu_count.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O.pytest_all.txt 2>&1; echo "pytest_all rc=$?" >> $O.pytest_all.txt
tail -5 $O.pytest_all.txt
