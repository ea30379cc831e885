#!/bin/bash
# round-2 multi-GPU batch (gpurun --gpus N): topology, multi-device tests, D2H ceiling, bench.py under torchrun
mkdir -p gpurun_out
N=${1:-2}
O=gpurun_out/r2m$N
{ nvidia-smi topo -m; nproc; cat /sys/fs/cgroup/cpuset.cpus.effective /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null; numactl -H 2>/dev/null | head -20;
  for d in /sys/bus/pci/devices/*; do if [ -f $d/numa_node ] && grep -q 0x10de $d/vendor 2>/dev/null; then echo "$d $(cat $d/numa_node) $(cat $d/class)"; fi; done; } > $O.topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O.pytest_multi.txt 2>&1; echo "rc=$?" >> $O.pytest_multi.txt
tail -5 $O.pytest_multi.txt
timeout 600 tools/_build/microbench_d2h 2 > $O.d2h.txt 2>&1; cat $O.d2h.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > $O.bench.json 2> $O.bench.err
tail -c 2500 $O.bench.json; tail -3 $O.bench.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $O.bench_ref.json 2>> $O.bench.err
