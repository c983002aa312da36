#!/bin/bash
O=gpurun_out/r2multi4; mkdir -p $O
timeout 600 python -m pytest tests/test_batched_gpu.py tests/test_device_nuts_gpu.py tests/test_stan_dropin_gpu.py -x -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log; tail -5 $O/tests.log
( time timeout 600 python bench.py --steps 20 --warmup 5 ) > $O/bench_default.json 2> $O/bench_default.err; echo "default rc=$?"
timeout 100 python bench_nuts.py --config 1 --chains 8 --driver device --ref-iters 0 > $O/nuts_cfg1_c8_device.json 2> $O/nuts_cfg1_c8_device.err
timeout 100 python bench_nuts.py --config 1 --chains 8 --ref-iters 0 > $O/nuts_cfg1_c8_service.json 2> $O/nuts_cfg1_c8_service.err
timeout 100 python bench_nuts.py --config 1 --chains 4 --driver device --ref-iters 0 > $O/nuts_cfg1_c4_device.json 2> $O/nuts_cfg1_c4_device.err
python - <<'PY'
import json,glob
O='gpurun_out/r2multi4'
try:
    d=json.loads(open(O+'/bench_default.json').read().strip().splitlines()[-1]); e=d.get('ess') or {}
    print('default', round(d['value'],2), round(d['e2e']['value'],1), d['roofline']['frac'])
    print('four', json.dumps(d.get('four_chains'))[:600])
    for k in ('b200','b200_device_driver'):
        x=e.get(k) or {}; print(k, {q: x.get(q) for q in ('wall_s','grad_evals_per_s','ess_min','ess_min_per_s','rounds')})
except Exception as ex: print('ERR', ex)
for f in sorted(glob.glob(O+'/nuts_*.json')):
    try:
        b=json.loads(open(f).read().strip().splitlines()[-1])['b200']; print(f.split('/')[-1], {k: b.get(k) for k in ('wall_s','grad_evals_per_s','ess_min_per_s')})
    except Exception as e: print(f,'ERR',e)
PY
grep real $O/bench_default.err
