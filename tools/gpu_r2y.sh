#!/bin/bash
# round-2 GPU batch Y: compute-sanitizer (memcheck, racecheck) over the second-generation kernels: long_kernel with the
# look-ahead ring (cp.async, three slots), MODE_FWD with replicas, bucket_kernel / count_kernel
mkdir -p gpurun_out
O=gpurun_out/r2y
SEL_MEM="k7_edge_lengths or k7_many_short or forward_fold_all_lengths or forward_fold_replicated or bucket_path_ragged or bucket_low_complexity or bucket_segment_sizes"
SEL_RACE="k7_edge_lengths or k7_many_short or forward_fold_replicated or bucket_low_complexity or bucket_segment_sizes"
{
echo "compute-sanitizer memcheck (-k \"$SEL_MEM\") and racecheck (-k \"$SEL_RACE\") over tests/test_gpu_long.py + tests/test_gpu_bucket.py on B200"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_long.py tests/test_gpu_bucket.py -m gpu -x -q -k "$SEL_MEM" 2>&1 | tail -8
echo "MEMCHECK_RC=${PIPESTATUS[0]}"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_long.py tests/test_gpu_bucket.py -m gpu -x -q -k "$SEL_RACE" > $O.racecheck_full.txt 2>&1; grep -o "in [a-z_]*\.cuh:[0-9]*" $O.racecheck_full.txt | sort | uniq -c | head -20; tail -5 $O.racecheck_full.txt
echo "RACECHECK_RC=$(grep -c "RACECHECK SUMMARY: 0 hazards" $O.racecheck_full.txt) (1 = clean)"
} > $O.sanitizer.txt 2>&1
cat $O.sanitizer.txt
