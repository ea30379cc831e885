#!/bin/bash
# round-2 last 2-GPU check of the committed tree: the multi-device tests (skipped on one-GPU boxes) and the gloo/NCCL rank test
mkdir -p gpurun_out
O=gpurun_out/r2g2
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_multi_rank.py -x -q > $O.pytest.txt 2>&1; echo "rc=$?" >> $O.pytest.txt
tail -4 $O.pytest.txt
