#!/bin/bash
O=gpurun_out/r2graph; mkdir -p $O
timeout 300 python -m pytest tests/test_device_nuts_gpu.py -x -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log; tail -4 $O/tests.log
for c in 4 8 64; do
  timeout 100 python bench_nuts.py --config 1 --chains $c --driver device --ref-iters 0 > $O/nuts_cfg1_c${c}_graph.json 2> $O/nuts_cfg1_c${c}_graph.err
  B200GLM_NO_GRAPH=1 timeout 100 python bench_nuts.py --config 1 --chains $c --driver device --ref-iters 0 > $O/nuts_cfg1_c${c}_nograph.json 2> $O/nuts_cfg1_c${c}_nograph.err
done
timeout 100 python bench_nuts.py --config 1 --chains 4 --ref-iters 0 > $O/nuts_cfg1_c4_service.json 2> $O/nuts_cfg1_c4_service.err
python - <<'PY'
import json,glob
O='gpurun_out/r2graph'
for f in sorted(glob.glob(O+'/nuts_*.json')):
    try:
        b=json.loads(open(f).read().strip().splitlines()[-1])['b200']; print(f.split('/')[-1], {k: b.get(k) for k in ('wall_s','grad_evals_per_s','ess_min_per_s','rounds')})
    except Exception as e: print(f,'ERR',e, open(f.replace('.json','.err')).read()[-500:])
PY
