#!/bin/bash
# round-2 GPU batch AH: full GPU suite on the final library + compute-sanitizer over the new paths (MODE_K8, ordering kernels)
mkdir -p gpurun_out
O=gpurun_out/r2ah
timeout 1800 python -m pytest tests -m gpu -x -q > $O.pytest_all.txt 2>&1; echo "pytest_all rc=$?" >> $O.pytest_all.txt
tail -4 $O.pytest_all.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O.smoke.txt 2>&1; tail -1 $O.smoke.txt
SEL="k8_long_kernel_mixed or longest_first_order or k7_edge_lengths or bucket_low_complexity"
{
echo "compute-sanitizer memcheck and racecheck (-k \"$SEL\") over tests/test_gpu_parity.py + tests/test_gpu_long.py + tests/test_gpu_bucket.py on B200 (final library)"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_long.py tests/test_gpu_bucket.py -m gpu -x -q -k "$SEL" 2>&1 | tail -6
echo "MEMCHECK_RC=${PIPESTATUS[0]}"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_long.py tests/test_gpu_bucket.py -m gpu -x -q -k "$SEL" > $O.racecheck_full.txt 2>&1
grep -o "in [a-z_]*\.cuh:[0-9]*" $O.racecheck_full.txt | sort | uniq -c | head -20; tail -4 $O.racecheck_full.txt
} > $O.sanitizer.txt 2>&1
cat $O.sanitizer.txt
