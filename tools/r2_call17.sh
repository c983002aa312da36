#!/bin/bash
O=gpurun_out/r2c17; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log; tail -6 $O/tests.log
timeout 200 python tools/timeline_probe.py --rows 1250000 --out $O/tl_n1_1250k.json > $O/tl.log 2>&1
timeout 200 python bench.py --rows 1250000 --steps 1000 --warmup 20 --no-cpu-baseline --no-parity --ess-iters 0 > $O/bench_1250k.json 2> $O/bench_1250k.err
python - <<'PY'
import json
O='gpurun_out/r2c17'
d=json.loads(open(f'{O}/tl_n1_1250k.json').readline()); print(d['us_per_step_events_plain'], {k:round(v,2) for k,v in d['phases_us'].items()}); print({k:round(v,2) for k,v in d['tail_fine_us'].items()})
d=json.loads(open(f'{O}/bench_1250k.json').read().strip().splitlines()[-1]); print('1250k', round(d['value'],1), round(d['ms_per_step'],5), round(d['e2e']['value'],1), round(d['e2e']['value']/d['value'],3))
PY
