#!/bin/bash
# round-2 GPU batch AK: default bench line with the file driver's 16 default threads (cli object, tmpfs leg), CLI tests
mkdir -p gpurun_out
O=gpurun_out/r2ak
timeout 900 python -m pytest tests/test_gpu_cli.py -m gpu -x -q > $O.pytest.txt 2>&1; tail -2 $O.pytest.txt
( time timeout 900 python bench.py ) > $O.bench_default.json 2> $O.bench_default.err; tail -3 $O.bench_default.err
python - <<PY
import json
d=json.loads(open('$O.bench_default.json').read().strip().splitlines()[-1])
c=d['cli']; print({k:(round(v,3) if isinstance(v,float) else v) for k,v in c.items() if k not in ('what','cpu_baseline','output_on_tmpfs')}); print('tmpfs', c.get('output_on_tmpfs'))
print(d['value'], d['e2e']['value'])
PY
