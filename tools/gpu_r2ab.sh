#!/bin/bash
# round-2 GPU batch AB: final validation of the committed kernels — full GPU suite, smoke, default bench line, reference
# arm, launch lists, and ncu --set full captures of the kernels that changed since batch W (bucket_kernel, contigs)
mkdir -p gpurun_out
O=gpurun_out/r2ab
KRE='regex:short_kernel|seq_kernel|long_kernel|bucket_kernel|count_kernel|tile_|order_|format_norm|rebase_offsets|wave_kernel|finalize'
timeout 1800 python -m pytest tests -m gpu -x -q > $O.pytest_all.txt 2>&1; echo "pytest_all rc=$?" >> $O.pytest_all.txt
tail -4 $O.pytest_all.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O.smoke.txt 2>&1; tail -2 $O.smoke.txt
( time timeout 900 python bench.py ) > $O.bench_default.json 2> $O.bench_default.err; tail -3 $O.bench_default.err
timeout 600 python bench.py --impl reference > $O.bench_reference.json 2> $O.bench_reference.err
LL="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none"
ncu $LL -k "$KRE" -c 12 --csv --log-file $O.launches_k5.csv python bench.py --scale 0.2 --steps 2 --warmup 3 --no-e2e --no-cpu --no-cli --no-per-config > $O.ncu_k5.log 2>&1
ncu $LL -k "$KRE" -c 6 --csv --log-file $O.launches_k7.csv python bench.py --workload reads10k_k7 --scale 0.2 --steps 2 --warmup 3 --no-e2e --no-cpu --no-cli --no-per-config > $O.ncu_k7.log 2>&1
ncu $LL -k "$KRE" -c 16 --csv --log-file $O.launches_contigs.csv python bench.py --workload contigs_k4 --steps 2 --warmup 3 --no-e2e --no-cpu --no-cli --no-per-config > $O.ncu_contigs.log 2>&1
ncu $LL -k "$KRE" -c 20 --csv --log-file $O.launches_k10.csv python bench.py --workload reads100k_k10 --steps 2 --warmup 3 --no-e2e --no-cpu --no-cli --no-per-config > $O.ncu_k10.log 2>&1
ncu $LL -k "$KRE" -c 8 --csv --log-file $O.launches_k8.csv python bench.py --workload reads10k_k8 --scale 0.3 --steps 2 --warmup 3 --no-e2e --no-cpu --no-cli --no-per-config > $O.ncu_k8.log 2>&1
prof() { # name regex workload scale
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -o $O.prof_$1 \
    python bench.py --workload $3 --scale $4 --steps 1 --warmup 3 --no-e2e --no-cpu --no-cli --no-per-config > $O.ncu_$1.log 2>&1
}
prof contigs long_kernel contigs_k4 0.25
prof bucket bucket_kernel reads100k_k10 0.5
du -sh gpurun_out
