#!/bin/bash
# full GPU suite + cfg5 at HBM scale with the lane-parallel wide-kernel producer + smoke
O=gpurun_out/r2full; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log; tail -6 $O/tests.log
timeout 400 python bench.py --config 5 --streamed --rows 0 --steps 10 --warmup 3 --no-cpu-baseline > $O/cfg5_full.json 2> $O/cfg5_full.err; echo "cfg5 rc=$?"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python - <<'PY'
import json
O='gpurun_out/r2full'
try:
    d=json.loads(open(O+'/cfg5_full.json').read().strip().splitlines()[-1])
    print('cfg5_full', d['config']['workload'], round(d['value'],2), round(d['ms_per_step'],3), round(d['e2e']['value'],2), d['roofline']['frac'], d['roofline']['achieved'], (d.get('parity') or {}).get('max_rel_err'))
except Exception as ex: print('ERR', ex)
PY
