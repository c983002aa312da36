#!/bin/bash
O=gpurun_out/r2wide; mkdir -p $O
W="timeout 120 python bench.py --config 5 --steps 50 --warmup 5 --no-cpu-baseline --no-parity"
$W --rows 1000000 > $O/wide_K1000.json 2> $O/wide_K1000.err
$W --rows 1000000 --cols 500 > $O/wide_K500.json 2> $O/wide_K500.err
$W --rows 1000000 --cols 300 > $O/wide_K300.json 2> $O/wide_K300.err
$W --rows 500000 --cols 2000 > $O/wide_K2000.json 2> $O/wide_K2000.err
timeout 400 python bench.py --config 5 --streamed --rows 0 --steps 10 --warmup 3 --no-cpu-baseline > $O/cfg5_full.json 2> $O/cfg5_full.err
timeout 900 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log; tail -5 $O/tests.log
python - <<'PY'
import json,glob
O='gpurun_out/r2wide'
for f in sorted(glob.glob(O+'/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], round(d['value'],2), round(d['ms_per_step'],5), round(d['e2e']['value'],2), d.get('roofline',{}).get('frac'), d.get('roofline',{}).get('achieved'), (d.get('parity') or {}).get('max_rel_err'))
    except Exception as e: print(f, 'ERR', e)
PY
