#!/bin/bash
# What the driver runs at round end, on one GPU: the full GPU suite, smoke(), the default bench line.
O=gpurun_out/final; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log; tail -5 $O/tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 400 python bench.py ) > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
O='gpurun_out/final'
try:
    d=json.loads(open(O+'/bench_default.json').read().strip().splitlines()[-1]); e=d.get('ess') or {}
    print('default', round(d['value'],2), round(d['e2e']['value'],1), d['roofline']['frac'], d['clocks'], (d.get('parity') or {}).get('ok'))
    print('four', (d.get('four_chains') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
    for k in ('b200','b200_device_driver'):
        x=e.get(k) or {}; print(k, {q: x.get(q) for q in ('wall_s','grad_evals_per_s','ess_min_per_s')})
except Exception as ex: print('ERR', ex); print(open(O+'/bench_default.err').read()[-1500:])
PY
grep real $O/bench_default.err
