#!/bin/bash
O=gpurun_out/r2c11; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log; tail -15 $O/tests.log
B="timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --ess-iters 0"
$B --family ordered_logistic --classes 5 > $O/ordlog_N10M_K100_C5.json 2> $O/ordlog.err
$B --family categorical_logit --classes 4 > $O/catlog_N10M_K100_C4.json 2> $O/catlog4.err
$B --family categorical_logit --classes 2 > $O/catlog_N10M_K100_C2.json 2> $O/catlog2.err
$B --family categorical_logit --classes 8 --cols 50 --rows 20000000 > $O/catlog_N20M_K50_C8.json 2> $O/catlog8.err
python - <<'PY'
import json,glob
O='gpurun_out/r2c11'
for f in sorted(glob.glob(O+'/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], round(d['value'],2), round(d['ms_per_step'],5), round(d['e2e']['value'],2), d.get('roofline',{}).get('frac'), d.get('roofline',{}).get('achieved'))
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 $O/ordlog.err $O/catlog4.err
