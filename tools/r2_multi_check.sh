#!/bin/bash
# GPU call of round 2, session 2 (d): few-chain FMA kernel (tests, cfg2 ESS arms), wide kernel with the ring-first policy
O=gpurun_out/r2multi; mkdir -p $O
timeout 600 python -m pytest tests/test_batched_gpu.py tests/test_device_nuts_gpu.py -x -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log; tail -15 $O/tests.log
W="timeout 120 python bench.py --config 5 --steps 50 --warmup 5 --no-cpu-baseline --no-parity"
$W --rows 1000000 > $O/wide_K1000.json 2> $O/wide_K1000.err
( time timeout 600 python bench.py --steps 20 --warmup 5 ) > $O/bench_default.json 2> $O/bench_default.err; echo "default rc=$?"
python - <<'PY'
import json,glob
O='gpurun_out/r2multi'
for f in sorted(glob.glob(O+'/wide*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], round(d['ms_per_step'],5), round(d['roofline']['frac'],4), round(d['e2e']['value'],1))
    except Exception as e: print(f, 'ERR', e)
try:
    d=json.loads(open(O+'/bench_default.json').read().strip().splitlines()[-1]); e=d.get('ess') or {}
    print('default', round(d['value'],2), round(d['e2e']['value'],1), d['roofline']['frac'])
    for k in ('b200','b200_device_driver'):
        print(k, json.dumps(e.get(k))[:700])
except Exception as ex: print('ERR', ex)
PY
grep real $O/bench_default.err
