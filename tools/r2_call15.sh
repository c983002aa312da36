#!/bin/bash
O=gpurun_out/r2c15; mkdir -p $O
timeout 200 python bench.py --rows 1250000 --steps 1000 --warmup 20 --no-cpu-baseline --no-parity --ess-iters 0 > $O/bench_1250k.json 2> $O/bench_1250k.err
( time timeout 560 python bench_nuts.py --config 3 --chains 1024 --warmup 300 --samples 100 ) > $O/nuts_cfg3_300_100.json 2> $O/nuts_cfg3_300_100.err; echo "nuts cfg3 rc=$?"
python - <<'PY'
import json
O='gpurun_out/r2c15'
d=json.loads(open(f'{O}/bench_1250k.json').read().strip().splitlines()[-1]); print('1250k', round(d['value'],1), round(d['ms_per_step'],5), round(d['e2e']['value'],1), round(d['e2e']['value']/d['value'],3))
PY
cut -c1-2600 $O/nuts_cfg3_300_100.json; tail -4 $O/nuts_cfg3_300_100.err
