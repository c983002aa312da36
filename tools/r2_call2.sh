#!/bin/bash
set -x
O=gpurun_out/r2c2; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
timeout 300 python tools/timeline_probe.py --rows 1250000 --out $O/tl_n1_1250k.json > $O/tl_n1_1250k.log 2>&1
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err
timeout 300 python bench.py --rows 1250000 --steps 500 --warmup 20 --no-cpu-baseline > $O/bench_n1_1250k.json 2> $O/bench_n1_1250k.err
B200GLM_NO_HOST_MIRROR=1 B200GLM_NO_INLINE_THETA=1 timeout 300 python bench.py --rows 1250000 --steps 500 --warmup 20 --no-cpu-baseline > $O/bench_n1_1250k_oldhost.json 2> $O/bench_n1_1250k_oldhost.err
tail -3 $O/tests.log; python - <<'PY'
import json
O='gpurun_out/r2c2'
d=json.loads(open(O+'/tl_n1_1250k.json').readline()); print(d['us_per_step_events_plain'], {k:round(v,2) for k,v in d['phases_us'].items()})
for f in ['bench_n1','bench_n1_1250k','bench_n1_1250k_oldhost']:
    try:
        d=json.loads(open(f'{O}/{f}.json').read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])
    except Exception as e: print(f, 'ERR', e)
PY
