#!/bin/bash
# round-2 final 8-GPU bench line (gpurun --gpus 8): bench.py under torchrun, weak scaling + the sharded k = 7 data set
mkdir -p gpurun_out
N=${1:-8}
O=gpurun_out/r2f$N
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > $O.bench.json 2> $O.bench.err
python - <<PY
import json
d=json.loads(open('$O.bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling')}, 'e2e', d['e2e']['value'], d['e2e'].get('pcie_frac'))
print({k:(v.get('gbases_s'), v.get('frac')) for k,v in d['roofline']['per_config'].items()})
PY
tail -2 $O.bench.err
