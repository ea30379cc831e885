#!/bin/bash
# round-2 GPU batch AG: captures for MODE_K8 (ncu --set full, launch list) and the refreshed default bench line
mkdir -p gpurun_out
O=gpurun_out/r2ag
KRE='regex:short_kernel|seq_kernel|long_kernel|bucket_kernel|count_kernel|tile_|order_|format_norm|rebase_offsets|wave_kernel|finalize'
( time timeout 900 python bench.py ) > $O.bench_default.json 2> $O.bench_default.err; tail -3 $O.bench_default.err
LL="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none"
ncu $LL -k "$KRE" -c 8 --csv --log-file $O.launches_k8.csv python bench.py --workload reads10k_k8 --scale 0.3 --steps 2 --warmup 3 --no-e2e --no-cpu --no-cli --no-per-config > $O.ncu_k8.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:long_kernel -s 2 -c 1 -o $O.prof_k8 \
  python bench.py --workload reads10k_k8 --scale 0.2 --steps 1 --warmup 3 --no-e2e --no-cpu --no-cli --no-per-config > $O.ncu_k8_full.log 2>&1
du -sh gpurun_out
