#!/bin/bash
O=gpurun_out/r2c12; mkdir -p $O
B="timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --ess-iters 0"
$B --family ordered_logistic --classes 5 > $O/ordlog_N10M_K100_C5.json 2> $O/ordlog.err
$B --family categorical_logit --classes 4 > $O/catlog_N10M_K100_C4.json 2> $O/catlog4.err
$B --family categorical_logit --classes 2 > $O/catlog_N10M_K100_C2.json 2> $O/catlog2.err
$B --family categorical_logit --classes 8 --cols 50 --rows 20000000 > $O/catlog_N20M_K50_C8.json 2> $O/catlog8.err
timeout 300 python -m pytest tests/test_class_models_gpu.py -m gpu -x -q 2>&1 | tail -2
for t in cfg2 cfg2shard cfg4shard wide cfg3 ordlog catlog; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:glm_ --launch-skip 4 --launch-count 2 -f -o $O/ncu_$t python tools/ncu_target.py $t > $O/ncu_$t.log 2>&1; echo "ncu $t rc=$?"
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_bench.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity --ess-iters 0 > $O/launches_bench.log 2>&1
( time timeout 900 python bench_nuts.py --config 3 --chains 1024 --warmup 300 --samples 100 ) > $O/nuts_cfg3.json 2> $O/nuts_cfg3.err; echo "nuts cfg3 rc=$?"
python - <<'PY'
import json,glob
O='gpurun_out/r2c12'
for f in sorted(glob.glob(O+'/*log*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], round(d['value'],2), round(d['ms_per_step'],5), round(d['e2e']['value'],2), d.get('roofline',{}).get('frac'), d.get('roofline',{}).get('achieved'))
    except Exception as e: print(f, 'ERR', e)
PY
cut -c1-2500 $O/nuts_cfg3.json; tail -5 $O/nuts_cfg3.err
ls -la $O/*.ncu-rep
