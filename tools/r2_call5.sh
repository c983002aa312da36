#!/bin/bash
set -x
O=gpurun_out/r2c5; mkdir -p $O
B="timeout 600 python bench.py --config 4 --steps 20 --warmup 3 --no-cpu-baseline --no-parity"
$B > $O/cfg4_50M_default.json 2> $O/cfg4_50M_default.err; echo "default rc=$?"
B200GLM_STAGES_MULT8=1 $B > $O/cfg4_50M_mult8.json 2> $O/cfg4_50M_mult8.err; echo "mult8 rc=$?"
B200GLM_PDL_PREFETCH=32 $B > $O/cfg4_50M_pf32.json 2> $O/cfg4_50M_pf32.err; echo "pf32 rc=$?"
B200GLM_NO_GROUP_FUSION=1 $B > $O/cfg4_50M_unfused.json 2> $O/cfg4_50M_unfused.err; echo "unfused rc=$?"
B200GLM_NO_PDL=1 $B > $O/cfg4_50M_nopdl.json 2> $O/cfg4_50M_nopdl.err; echo "nopdl rc=$?"
$B --rows 20000000 > $O/cfg4_20M_default.json 2> $O/cfg4_20M_default.err; echo "20M rc=$?"
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python bench.py --config 4 --steps 2 --warmup 3 --no-cpu-baseline --no-parity > $O/sanitizer.log 2>&1; echo "sanitizer rc=$?"
grep -E "Invalid|Error|error|at 0x|by thread|Address" $O/sanitizer.log | head -30
timeout 900 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log; tail -3 $O/tests.log
timeout 300 python tools/timeline_probe.py --rows 1250000 --out $O/tl_n1_1250k.json > $O/tl.log 2>&1
python - <<'PY'
import json,glob
O='gpurun_out/r2c5'
d=json.loads(open(f'{O}/tl_n1_1250k.json').readline()); print(d['us_per_step_events_plain'], {k:round(v,2) for k,v in d['phases_us'].items()}); print({k:round(v,2) for k,v in d['tail_fine_us'].items()})
for f in sorted(glob.glob(O+'/cfg4*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], round(d['value'],1), round(d['ms_per_step'],5), round(d['e2e']['value'],1), d.get('roofline',{}).get('frac'), d.get('gpu_launches'))
    except Exception as e: print(f, 'ERR', e)
PY
