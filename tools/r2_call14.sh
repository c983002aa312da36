#!/bin/bash
O=gpurun_out/r2c14; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log; tail -4 $O/tests.log
B="timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --ess-iters 0 --no-parity"
$B --config 5 --rows 1000000 > $O/wide_K1000.json 2> $O/wide_K1000.err
$B --config 5 --rows 1000000 --cols 500 > $O/wide_K500.json 2> $O/wide_K500.err
$B --family ordered_logistic --classes 5 > $O/ordlog_N10M_K100_C5.json 2> $O/ordlog.err
$B --family categorical_logit --classes 4 > $O/catlog_N10M_K100_C4.json 2> $O/catlog4.err
$B --family categorical_logit --classes 2 > $O/catlog_N10M_K100_C2.json 2> $O/catlog2.err
$B --family categorical_logit --classes 8 --cols 50 --rows 20000000 > $O/catlog_N20M_K50_C8.json 2> $O/catlog8.err
timeout 200 python tools/timeline_probe.py --rows 1250000 --out $O/tl_n1_1250k.json > $O/tl.log 2>&1
python - <<'PY'
import json,glob
O='gpurun_out/r2c14'
try:
    d=json.loads(open(f'{O}/tl_n1_1250k.json').readline()); print(d['us_per_step_events_plain'], {k:round(v,2) for k,v in d['phases_us'].items()})
except Exception as e: print('tl ERR', e)
for f in sorted(glob.glob(O+'/*.json')):
    if 'tl_' in f: continue
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], round(d['value'],2), round(d['ms_per_step'],5), round(d['e2e']['value'],2), d.get('roofline',{}).get('frac'), d.get('roofline',{}).get('achieved'))
    except Exception as e: print(f, 'ERR', e)
PY
