#!/bin/bash
set -x
O=gpurun_out/r2c4; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
for D in 0 1 2 3 4 8; do
  B200GLM_PDL_PREFETCH=$D timeout 300 python bench.py --rows 1250000 --steps 1000 --warmup 20 --no-cpu-baseline --no-parity --ess-iters 0 > $O/bench_1250k_D$D.json 2> $O/bench_1250k_D$D.err
done
B200GLM_PDL_PREFETCH=2 timeout 300 python tools/timeline_probe.py --rows 1250000 --out $O/tl_n1_1250k_D2.json > $O/tl_D2.log 2>&1
B200GLM_PDL_PREFETCH=0 timeout 300 python tools/timeline_probe.py --rows 1250000 --out $O/tl_n1_1250k_D0.json > $O/tl_D0.log 2>&1
timeout 600 python bench.py --config 4 --rows 5000000 --steps 100 --warmup 10 --no-cpu-baseline --no-parity > $O/bench_cfg4_5M.json 2> $O/bench_cfg4_5M.err
B200GLM_STAGES_MULT8=1 timeout 600 python bench.py --config 4 --rows 5000000 --steps 100 --warmup 10 --no-cpu-baseline --no-parity > $O/bench_cfg4_5M_mult8.json 2> $O/bench_cfg4_5M_mult8.err
timeout 600 python bench.py --config 4 --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_cfg4_50M.json 2> $O/bench_cfg4_50M.err
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $O/bench_default.json 2> $O/bench_default.err
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > $O/bench_ref.json 2> $O/bench_ref.err
tail -3 $O/tests.log; python - <<'PY'
import json,glob
O='gpurun_out/r2c4'
for f in ['tl_n1_1250k_D2','tl_n1_1250k_D0']:
    try:
        d=json.loads(open(f'{O}/{f}.json').readline()); print(f, d['us_per_step_events_plain'], {k:round(v,2) for k,v in d['phases_us'].items()}); print({k:round(v,2) for k,v in d['tail_fine_us'].items()})
    except Exception as e: print(f,'ERR',e)
for f in sorted(glob.glob(O+'/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], round(d['value'],1), round(d['ms_per_step'],5), round(d['e2e']['value'],1), d.get('roofline',{}).get('frac'), d.get('gpu_launches'))
    except Exception as e: print(f, 'ERR', e)
PY
tail -5 $O/bench_default.err; tail -5 $O/bench_ref.err; cut -c1-3000 $O/bench_default.json
