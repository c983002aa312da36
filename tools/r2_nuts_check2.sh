#!/bin/bash
# GPU call of round 2, session 2 (b): wide-kernel producer forms, device-side NUTS at config 3 / config 1
O=gpurun_out/r2nuts2; mkdir -p $O
W="timeout 120 python bench.py --config 5 --steps 50 --warmup 5 --no-cpu-baseline --no-parity"
for mode in single lanes poll; do
  B200GLM_WIDE_PRODUCER=$mode $W --rows 1000000 > $O/wide_K1000_$mode.json 2> $O/wide_K1000_$mode.err
  B200GLM_WIDE_PRODUCER=$mode $W --rows 1000000 --cols 500 > $O/wide_K500_$mode.json 2> $O/wide_K500_$mode.err
  B200GLM_WIDE_PRODUCER=$mode $W --rows 500000 --cols 2000 > $O/wide_K2000_$mode.json 2> $O/wide_K2000_$mode.err
  B200GLM_WIDE_PRODUCER=$mode $W --rows 1000000 --cols 300 > $O/wide_K300_$mode.json 2> $O/wide_K300_$mode.err
done
timeout 300 python -m pytest tests/test_device_nuts_gpu.py -x -q > $O/tests_nuts.log 2>&1; echo "tests rc=$?" >> $O/tests_nuts.log; tail -5 $O/tests_nuts.log
timeout 500 python bench_nuts.py --config 3 --rows 100000 --chains 1024 --warmup 400 --samples 100 --driver device > $O/nuts_cfg3_100k_device.json 2> $O/nuts_cfg3_100k_device.err
timeout 100 python bench_nuts.py --config 1 --driver device --ref-iters 0 > $O/nuts_cfg1_device.json 2> $O/nuts_cfg1_device.err
timeout 100 python bench_nuts.py --config 1 --chains 64 --driver device --ref-iters 0 > $O/nuts_cfg1_c64_device.json 2> $O/nuts_cfg1_c64_device.err
timeout 100 python bench_nuts.py --config 1 --chains 64 --ref-iters 0 > $O/nuts_cfg1_c64_service.json 2> $O/nuts_cfg1_c64_service.err
timeout 100 python bench_nuts.py --config 1 --chains 64 --driver batched --ref-iters 0 > $O/nuts_cfg1_c64_batched.json 2> $O/nuts_cfg1_c64_batched.err
python - <<'PY'
import json,glob
O='gpurun_out/r2nuts2'
for f in sorted(glob.glob(O+'/wide*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], round(d['ms_per_step'],5), round(d['roofline']['frac'],4))
    except Exception as e: print(f, 'ERR', e)
for f in sorted(glob.glob(O+'/nuts_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])['b200']; print(f.split('/')[-1], {k: d.get(k) for k in ('wall_s','grad_evals_per_s','ess_min','ess_min_per_s','divergent','mean_n_leapfrog','rounds','lanes','ols_check')})
    except Exception as e: print(f, 'ERR', e)
PY
