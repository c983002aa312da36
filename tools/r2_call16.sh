#!/bin/bash
O=gpurun_out/r2c16; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log; tail -6 $O/tests.log
( time timeout 600 python bench.py --steps 20 --warmup 5 ) > $O/bench_default.json 2> $O/bench_default.err; echo "default rc=$?"
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 ) > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python - <<'PY'
import json
O='gpurun_out/r2c16'
for f in ['bench_default','bench_ref']:
    try:
        d=json.loads(open(f'{O}/{f}.json').read().strip().splitlines()[-1]); print(f, round(d['value'],2), round(d['ms_per_step'],4), d['e2e']['value'], d.get('roofline',{}).get('frac'), (d.get('parity') or {}).get('max_rel_err'), ((d.get('ess') or {}).get('b200') or {}).get('ess_min_per_s'), d.get('cpu_baseline',{}).get('sample'))
    except Exception as e: print(f,'ERR',e)
PY
grep real $O/bench_default.err $O/bench_ref.err
