#!/bin/bash
set -x
O=gpurun_out/r2c3; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
timeout 300 python tools/timeline_probe.py --rows 1250000 --out $O/tl_n1_1250k.json > $O/tl_n1_1250k.log 2>&1
B200GLM_TL_REPEAT=1 timeout 300 python tools/timeline_probe.py --rows 1250000 --out $O/tl_n1_1250k_repeat.json > $O/tl_n1_1250k_repeat.log 2>&1
timeout 600 python bench.py --config 4 --rows 5000000 --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_cfg4_5M.json 2> $O/bench_cfg4_5M.err
B200GLM_NO_GROUP_FUSION=1 timeout 600 python bench.py --config 4 --rows 5000000 --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_cfg4_5M_unfused.json 2> $O/bench_cfg4_5M_unfused.err
timeout 600 python bench.py --config 4 --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_cfg4_50M.json 2> $O/bench_cfg4_50M.err
tail -3 $O/tests.log; python - <<'PY'
import json
O='gpurun_out/r2c3'
for f in ['tl_n1_1250k','tl_n1_1250k_repeat']:
    d=json.loads(open(f'{O}/{f}.json').readline()); print(d['us_per_step_events_plain'], {k:round(v,2) for k,v in d['phases_us'].items()}); print({k:round(v,2) for k,v in d['tail_fine_us'].items()}, d['ns_since_first_cta_entry'].get('fine_cold_pass_done'), d['ns_since_first_cta_entry']['fine_fence_after_ticket'], d['ns_since_first_cta_entry']['fine_sums_written'])
for f in ['bench_cfg4_5M','bench_cfg4_5M_unfused','bench_cfg4_50M']:
    try:
        d=json.loads(open(f'{O}/{f}.json').read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'])
    except Exception as e: print(f, 'ERR', e)
PY
