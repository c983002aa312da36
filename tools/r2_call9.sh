#!/bin/bash
O=gpurun_out/r2c9; mkdir -p $O
timeout 600 python -m pytest tests/test_batched_gpu.py tests/test_parity_gpu.py -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log; tail -4 $O/tests.log
W="timeout 300 python bench.py --config 5 --steps 50 --warmup 5 --no-cpu-baseline --no-parity"
B200GLM_WIDE_ROWS=4 $W --rows 1000000 > $O/wide_K1000_wr4.json 2> $O/wide_K1000_wr4.err
B200GLM_WIDE_ROWS=8 $W --rows 1000000 > $O/wide_K1000_wr8.json 2> $O/wide_K1000_wr8.err
B200GLM_WIDE_ROWS=16 $W --rows 1000000 --cols 500 > $O/wide_K500_wr16.json 2> $O/wide_K500_wr16.err
B200GLM_WIDE_ROWS=8 $W --rows 1000000 --cols 500 > $O/wide_K500_wr8.json 2> $O/wide_K500_wr8.err
B200GLM_WIDE_ROWS=4 $W --rows 1000000 --cols 500 > $O/wide_K500_wr4.json 2> $O/wide_K500_wr4.err
B200GLM_WIDE_ROWS=4 $W --rows 500000 --cols 2000 > $O/wide_K2000_wr4.json 2> $O/wide_K2000_wr4.err
timeout 600 python bench.py --config 3 --steps 20 --warmup 3 --no-cpu-baseline > $O/cfg3.json 2> $O/cfg3.err; echo "cfg3 rc=$?"
python - <<'PY'
import json,glob
O='gpurun_out/r2c9'
for f in sorted(glob.glob(O+'/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], round(d['value'],2), round(d['ms_per_step'],5), round(d['e2e']['value'],2), d.get('roofline',{}).get('frac'), d.get('roofline',{}).get('achieved'), d.get('roofline',{}).get('peak'))
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 $O/cfg3.err
