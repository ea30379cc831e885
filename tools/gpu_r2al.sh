#!/bin/bash
# round-2 GPU batch AL: last validation of the committed tree — full GPU suite, smoke, default bench line, reference arm
mkdir -p gpurun_out
O=gpurun_out/r2al
timeout 1800 python -m pytest tests -m gpu -x -q > $O.pytest_all.txt 2>&1; echo "pytest_all rc=$?" >> $O.pytest_all.txt
tail -4 $O.pytest_all.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O.smoke.txt 2>&1; tail -2 $O.smoke.txt
( time timeout 900 python bench.py ) > $O.bench_default.json 2> $O.bench_default.err; tail -3 $O.bench_default.err
timeout 600 python bench.py --impl reference > $O.bench_reference.json 2> $O.bench_reference.err
python - <<PY
import json
d=json.loads(open('$O.bench_default.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, d['roofline']['frac'])
for k,v in d['roofline']['per_config'].items(): print(k, round(v['gbases_s'],1), round(v['frac'],3), v['parity_exact'])
print('e2e', d['e2e']['value'], d['e2e']['pcie_frac'], 'cpu', d['cpu_baseline']['value'])
c=d['cli']; print('cli', c['text_gb_per_s'], c['total_ms'], c['write_frac'], c['output_on_tmpfs'])
r=json.loads(open('$O.bench_reference.json').read().strip().splitlines()[-1]); print('ref', r['value'], r['impl'])
PY
