#!/bin/bash
O=gpurun_out/r2multi3; mkdir -p $O
timeout 600 python -m pytest tests/test_batched_gpu.py tests/test_device_nuts_gpu.py -x -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log; tail -5 $O/tests.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:glm_multi --launch-skip 2 --launch-count 1 -f -o $O/multi4 python tools/ncu_target.py multi4 > $O/ncu.log 2>&1; echo "ncu rc=$?"
NCU_SUMMARY_DIR=$O python profiles/summarize_ncu.py $O/multi4.ncu-rep r2_glm_multi_bernoulli_N4M_K100_C4 4000000 100 bernoulli_logit 0 glm_multi > $O/sum.log 2>&1
grep -E "gpu__time_duration|dram__bytes_read|issue_active|pipe_fp64|short_scoreboard|stalled_wait|long_scoreboard|branch_resolving|registers" $O/r2_glm_multi_bernoulli_N4M_K100_C4_summary.txt | cut -c1-130
( time timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ) > $O/bench_default.json 2> $O/bench_default.err; echo "default rc=$?"
python - <<'PY'
import json
O='gpurun_out/r2multi3'
try:
    d=json.loads(open(O+'/bench_default.json').read().strip().splitlines()[-1]); e=d.get('ess') or {}
    print('default', round(d['value'],2), round(d['e2e']['value'],1), d['roofline']['frac'])
    for k in ('b200','b200_device_driver'):
        x=e.get(k) or {}; print(k, {q: x.get(q) for q in ('wall_s','grad_evals_per_s','ess_min','ess_min_per_s','rounds')})
except Exception as ex: print('ERR', ex)
PY
