#!/bin/bash
# round-2 GPU batch J: replicas for MODE_FWD, SIMD slow decode path; final ncu captures and the CPU arm of every config
mkdir -p gpurun_out
O=gpurun_out/r2j
timeout 1800 python -m pytest tests -m gpu -x -q > $O.pytest_all.txt 2>&1; echo "pytest_all rc=$?" >> $O.pytest_all.txt
tail -5 $O.pytest_all.txt
run() { # workload scale opts...
  w=$1; sc=$2; shift 2; flags=""; for kv in "$@"; do flags="$flags --opt $kv"; done
  timeout 300 python bench.py --workload $w --scale $sc --steps 5 --no-e2e --no-cpu --no-cli --no-per-config $flags 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$w', '$sc', '$*', 'Gb/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'rows1', d['rows_sum_to_one'])"
}
{
run contigs_k4 1.0 fwd_replicas=0
run contigs_k4 1.0 fwd_replicas=1
run reads10k_k5 0.3
run reads10k_k7 1.0
run reads10k_k8 1.0
run reads100k_k10 1.0
run reads150_k5 1.0
run reads150_k3 1.0
run reads150_k4 1.0
run reads150_k6 1.0
run reads150_k7 1.0
} > $O.sweep.txt 2>&1
cat $O.sweep.txt
for w in reads10k_k7 contigs_k4 reads10k_k8 reads100k_k10; do
  python bench.py --impl reference --workload $w --steps 3 --warmup 1 | tail -1 >> $O.reference_arms.jsonl
done
python -c "
import json
for l in open('$O.reference_arms.jsonl'): d=json.loads(l); print(d['config']['workload'], round(d['value'],3), d['cpu_baseline']['cores'])"
prof() { # name regex workload scale
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -o $O.prof_$1 \
    python bench.py --workload $3 --scale $4 --steps 1 --warmup 3 --no-e2e --no-cpu --no-cli --no-per-config > $O.ncu_$1.log 2>&1
}
prof k7 long_kernel reads10k_k7 0.05
prof contigs long_kernel contigs_k4 0.25
prof bucket bucket_kernel reads100k_k10 0.5
prof count count_kernel reads100k_k10 0.5
prof short short_kernel reads150_k5 0.2
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"ktb" -c 12 --csv --log-file $O.launches_k5.csv \
  python bench.py --scale 0.2 --steps 2 --warmup 3 --no-e2e --no-cpu --no-cli --no-per-config > $O.ncu_k5.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"ktb" -c 12 --csv --log-file $O.launches_k10.csv \
  python bench.py --workload reads100k_k10 --steps 2 --warmup 3 --no-e2e --no-cpu --no-cli --no-per-config > $O.ncu_k10.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"ktb" -c 6 --csv --log-file $O.launches_k7.csv \
  python bench.py --workload reads10k_k7 --scale 0.2 --steps 2 --warmup 3 --no-e2e --no-cpu --no-cli --no-per-config > $O.ncu_k7.log 2>&1
