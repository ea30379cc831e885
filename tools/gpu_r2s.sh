#!/bin/bash
# round-2 GPU batch S: ncu captures of the changed kernels (k = 7 with look-ahead, count_kernel), plus count_kernel's
# skeleton (probe build, every phase off) to see where an item's latency goes
mkdir -p gpurun_out
O=gpurun_out/r2s
prof() { # name regex workload scale
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -o $O.prof_$1 \
    python bench.py --workload $3 --scale $4 --steps 1 --warmup 3 --no-e2e --no-cpu --no-cli --no-per-config > $O.ncu_$1.log 2>&1
}
prof k7 long_kernel reads10k_k7 0.05
prof count count_kernel reads100k_k10 0.5
cp tools/_probe/libkmertools_b200.so kmertools_b200/lib/libkmertools_b200.so
export KTB_COUNT_PROBE=15
prof count_skeleton count_kernel reads100k_k10 0.5
du -sh gpurun_out
