#!/bin/bash
# GPU call of round 2, session 2 (c): wide-kernel ring depth at K = 1000 (chain state on chip vs a deeper ring), default bench line
O=gpurun_out/r2nuts3; mkdir -p $O
W="timeout 120 python bench.py --config 5 --steps 50 --warmup 5 --no-cpu-baseline --no-parity --rows 1000000"
for mode in single lanes poll; do
  B200GLM_NO_STATE_SMEM=1 B200GLM_WIDE_PRODUCER=$mode $W > $O/wide_K1000_nostate_$mode.json 2> $O/wide_K1000_nostate_$mode.err
done
B200GLM_NO_STATE_SMEM=1 B200GLM_WIDE_PRODUCER=lanes $W --cols 500 > $O/wide_K500_nostate_lanes.json 2> $O/wide_K500_nostate_lanes.err
( time timeout 600 python bench.py --steps 20 --warmup 5 ) > $O/bench_default.json 2> $O/bench_default.err; echo "default rc=$?"
timeout 420 python bench_nuts.py --config 3 --chains 1024 --warmup 150 --samples 100 --driver device > $O/nuts_cfg3_1M_device.json 2> $O/nuts_cfg3_1M_device.err; echo "cfg3 rc=$?"
python - <<'PY'
import json,glob
O='gpurun_out/r2nuts3'
for f in sorted(glob.glob(O+'/wide*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], round(d['ms_per_step'],5), round(d['roofline']['frac'],4), round(d['e2e']['value'],1))
    except Exception as e: print(f, 'ERR', e)
try:
    d=json.loads(open(O+'/bench_default.json').read().strip().splitlines()[-1]); e=d.get('ess') or {}
    print('default', round(d['value'],2), round(d['e2e']['value'],1), d['roofline']['frac'], (d.get('parity') or {}))
    for k in ('b200','b200_device_driver'):
        print(k, json.dumps(e.get(k))[:700])
except Exception as ex: print('ERR', ex)
PY
grep real $O/bench_default.err
python -c "import json; d=json.load(open('$O/nuts_cfg3_1M_device.json'))['b200']; print({k: d.get(k) for k in ('wall_s','grad_evals_per_s','ess_min','ess_min_per_s','divergent','rounds','lanes','mean_lanes_per_round')})"
