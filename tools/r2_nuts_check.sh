#!/bin/bash
# GPU call of round 2, session 2: device-side NUTS tests + config 1 / config 3 runs, wide-kernel producer A/B
O=gpurun_out/r2nuts; mkdir -p $O
W="timeout 120 python bench.py --config 5 --steps 50 --warmup 5 --no-cpu-baseline --no-parity --rows 1000000"
B200GLM_WIDE_PRODUCER=single $W > $O/wide_K1000_single.json 2> $O/wide_K1000_single.err
B200GLM_WIDE_PRODUCER=lanes $W > $O/wide_K1000_lanes.json 2> $O/wide_K1000_lanes.err
B200GLM_WIDE_PRODUCER=single $W --cols 500 > $O/wide_K500_single.json 2> $O/wide_K500_single.err
B200GLM_WIDE_PRODUCER=lanes $W --cols 500 > $O/wide_K500_lanes.json 2> $O/wide_K500_lanes.err
timeout 600 python -m pytest tests/test_device_nuts_gpu.py -x -q > $O/tests_nuts.log 2>&1; echo "tests rc=$?" >> $O/tests_nuts.log; tail -25 $O/tests_nuts.log
timeout 200 python bench_nuts.py --config 1 --driver device --ref-iters 0 > $O/nuts_cfg1_device.json 2> $O/nuts_cfg1_device.err
timeout 200 python bench_nuts.py --config 1 --ref-iters 0 > $O/nuts_cfg1_service.json 2> $O/nuts_cfg1_service.err
python - <<'PY'
import json,glob
O='gpurun_out/r2nuts'
for f in sorted(glob.glob(O+'/wide*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], round(d['ms_per_step'],5), d['roofline']['frac'])
    except Exception as e: print(f, 'ERR', e)
for f in sorted(glob.glob(O+'/nuts_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])['b200']; print(f.split('/')[-1], {k: d[k] for k in ('wall_s','grad_evals_per_s','ess_min','ess_min_per_s','divergent','mean_n_leapfrog')})
    except Exception as e: print(f, 'ERR', e)
PY
