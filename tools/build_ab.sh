#!/bin/bash
# Builds tools/_probe/libA.so from the sources of the last commit (git HEAD), next to the working tree's library, so that a
# GPU batch can time "A = previous commit" against "B = working tree" on ONE box (box-to-box noise is +-3 %, same-box
# repeats agree to 0.1 %).  The gpu_r2t.sh ... gpu_r2af.sh batches copy one or the other over
# kmertools_b200/lib/libkmertools_b200.so on the (ephemeral) GPU box before each timed run.
set -e
cd "$(dirname "$0")/.."
rm -rf /tmp/headsrc && mkdir -p /tmp/headsrc tools/_probe
git archive HEAD kmertools_b200/csrc include | tar -x -C /tmp/headsrc
( cd /tmp/headsrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O3 -ccbin /usr/bin/g++ \
    -cudart static --shared $KTB_NVCC_EXTRA -o "$OLDPWD/tools/_probe/libA.so" kmertools_b200/csrc/*.cu kmertools_b200/csrc/*.cpp -lz )
python -m kmertools_b200.build
ls -la tools/_probe/libA.so kmertools_b200/lib/libkmertools_b200.so
