#!/bin/bash
# round-2 GPU batch AN: ncu --set full captures of the final binary (k = 7, count_kernel) so that the summaries under
# profiles/ belong to the committed tree
mkdir -p gpurun_out
O=gpurun_out/r2an
prof() { # name regex workload scale
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -o $O.prof_$1 \
    python bench.py --workload $3 --scale $4 --steps 1 --warmup 3 --no-e2e --no-cpu --no-cli --no-per-config > $O.ncu_$1.log 2>&1
}
prof k7 long_kernel reads10k_k7 0.05
prof count count_kernel reads100k_k10 0.5
du -sh gpurun_out
