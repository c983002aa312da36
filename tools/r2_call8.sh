#!/bin/bash
O=gpurun_out/r2c8; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log; tail -5 $O/tests.log
timeout 300 python bench.py --config 5 --rows 1000000 --steps 50 --warmup 5 --no-cpu-baseline > $O/wide_1M_K1000.json 2> $O/wide_1M_K1000.err; echo "wide rc=$?"
timeout 300 python bench.py --config 5 --rows 1000000 --cols 500 --steps 50 --warmup 5 --no-cpu-baseline > $O/wide_1M_K500.json 2> $O/wide_1M_K500.err
timeout 300 python bench.py --config 5 --rows 500000 --cols 2000 --steps 50 --warmup 5 --no-cpu-baseline > $O/wide_500k_K2000.json 2> $O/wide_500k_K2000.err
( time timeout 900 python bench.py --config 5 --streamed --rows 0 --steps 10 --warmup 3 --no-cpu-baseline ) > $O/cfg5_full.json 2> $O/cfg5_full.err; echo "cfg5 full rc=$?"
timeout 200 python tools/timeline_probe.py --rows 1250000 --out $O/tl_n1_1250k.json > $O/tl.log 2>&1
python - <<'PY'
import json,glob
O='gpurun_out/r2c8'
try:
    d=json.loads(open(f'{O}/tl_n1_1250k.json').readline()); print(d['us_per_step_events_plain'], {k:round(v,2) for k,v in d['phases_us'].items()}); print({k:round(v,2) for k,v in d['tail_fine_us'].items()})
except Exception as e: print('tl ERR', e)
for f in sorted(glob.glob(O+'/*.json')):
    if 'tl_' in f: continue
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], round(d['value'],2), round(d['ms_per_step'],5), round(d['e2e']['value'],2), d.get('roofline',{}).get('frac'), d.get('roofline',{}).get('frac_of_read_only_stream'), d['config'].get('rows_per_gpu'), d['config'].get('streamed_build_s'), (d.get('parity') or {}).get('max_rel_err'))
    except Exception as e: print(f, 'ERR', e)
PY
tail -4 $O/cfg5_full.err
