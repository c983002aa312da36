#!/bin/bash
O=gpurun_out/r2c7; mkdir -p $O
PROBE_STAGES=8,9,10,12,13,15,16 timeout 150 python tools/stage_probe.py 20000000 50 0 > $O/p1.log 2>&1; tail -8 $O/p1.log
PROBE_STAGES=13,11,9 timeout 150 python tools/stage_probe.py 20000000 50 1000 > $O/p3.log 2>&1; tail -4 $O/p3.log
PROBE_STAGES=8,7,6,5 timeout 150 python tools/stage_probe.py 10000000 100 0 bernoulli_logit > $O/p6.log 2>&1; tail -4 $O/p6.log
timeout 600 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log; tail -3 $O/tests.log
timeout 200 python tools/timeline_probe.py --rows 1250000 --out $O/tl_n1_1250k.json > $O/tl.log 2>&1
B="timeout 300 python bench.py --config 4 --no-cpu-baseline"
$B --steps 20 --warmup 3 > $O/cfg4_50M.json 2> $O/cfg4_50M.err; echo "cfg4 50M rc=$?"
$B --rows 6250000 --steps 100 --warmup 10 > $O/cfg4_6250k.json 2> $O/cfg4_6250k.err; echo "cfg4 6.25M rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err; echo "default rc=$?"
timeout 300 python bench.py --rows 1250000 --steps 1000 --warmup 20 --no-cpu-baseline --no-parity --ess-iters 0 > $O/bench_1250k.json 2> $O/bench_1250k.err
python - <<'PY'
import json,glob
O='gpurun_out/r2c7'
try:
    d=json.loads(open(f'{O}/tl_n1_1250k.json').readline()); print(d['us_per_step_events_plain'], {k:round(v,2) for k,v in d['phases_us'].items()}); print({k:round(v,2) for k,v in d['tail_fine_us'].items()})
except Exception as e: print('tl ERR', e)
for f in sorted(glob.glob(O+'/*.json')):
    if 'tl_' in f: continue
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], round(d['value'],1), round(d['ms_per_step'],5), round(d['e2e']['value'],1), d.get('roofline',{}).get('frac'), d.get('gpu_launches'), d.get('clocks'), (d.get('parity') or {}).get('max_rel_err'))
    except Exception as e: print(f, 'ERR', e)
PY
