#!/bin/bash
# round-2 GPU batch K: final ncu captures, part $1 (kept under the 64 MiB gpurun_out limit)
mkdir -p gpurun_out
O=gpurun_out/r2k
prof() { # name regex workload scale
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -o $O.prof_$1 \
    python bench.py --workload $3 --scale $4 --steps 1 --warmup 3 --no-e2e --no-cpu --no-cli --no-per-config > $O.ncu_$1.log 2>&1
}
if [ "$1" = "1" ]; then
  prof k7 long_kernel reads10k_k7 0.05
  prof contigs long_kernel contigs_k4 0.25
  prof short short_kernel reads150_k5 0.2
else
  prof bucket bucket_kernel reads100k_k10 0.5
  prof count count_kernel reads100k_k10 0.5
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"ktb" -c 12 --csv --log-file $O.launches_k5.csv \
    python bench.py --scale 0.2 --steps 2 --warmup 3 --no-e2e --no-cpu --no-cli --no-per-config > $O.ncu_k5.log 2>&1
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"ktb" -c 12 --csv --log-file $O.launches_k10.csv \
    python bench.py --workload reads100k_k10 --steps 2 --warmup 3 --no-e2e --no-cpu --no-cli --no-per-config > $O.ncu_k10.log 2>&1
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"ktb" -c 6 --csv --log-file $O.launches_k7.csv \
    python bench.py --workload reads10k_k7 --scale 0.2 --steps 2 --warmup 3 --no-e2e --no-cpu --no-cli --no-per-config > $O.ncu_k7.log 2>&1
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"ktb" -c 6 --csv --log-file $O.launches_contigs.csv \
    python bench.py --workload contigs_k4 --steps 2 --warmup 3 --no-e2e --no-cpu --no-cli --no-per-config > $O.ncu_contigs.log 2>&1
fi
du -sh gpurun_out
