#!/bin/bash
# 2-GPU box: chains sharded over GPUs (bench.py --config 3, bench_nuts --config 3 --driver device), 2-rank parity tests, device NUTS tests
O=gpurun_out/r2two; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_device_nuts_gpu.py tests/test_multigpu_gpu.py -x -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log; tail -4 $O/tests.log
timeout 200 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline > $O/cfg3_n1.json 2> $O/cfg3_n1.err; echo "cfg3 n1 rc=$?"
timeout 200 $TR --master-port 29911 bench.py --config 3 --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > $O/cfg3_n2.json 2> $O/cfg3_n2.err; echo "cfg3 n2 rc=$?"
timeout 300 $TR --master-port 29912 bench_nuts.py --config 3 --rows 100000 --chains 256 --warmup 100 --samples 50 --driver device > $O/nuts_cfg3_n2.json 2> $O/nuts_cfg3_n2.err; echo "nuts n2 rc=$?"
timeout 300 python bench_nuts.py --config 3 --rows 100000 --chains 256 --warmup 100 --samples 50 --driver device > $O/nuts_cfg3_n1.json 2> $O/nuts_cfg3_n1.err; echo "nuts n1 rc=$?"
python - <<'PY'
import json
O='gpurun_out/r2two'
for f in ('cfg3_n1','cfg3_n2'):
    try:
        d=json.loads(open(f'{O}/{f}.json').read().strip().splitlines()[-1]); print(f, d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), round(d['roofline']['frac'],3), d['config'].get('sharding'))
    except Exception as e: print(f,'ERR',e); print(open(f'{O}/{f}.err').read()[-800:])
for f in ('nuts_cfg3_n1','nuts_cfg3_n2'):
    try:
        d=json.loads(open(f'{O}/{f}.json').read().strip().splitlines()[-1]); b=d['b200']; print(f, d.get('n_gpus'), {k:b.get(k) for k in ('wall_s','grad_evals_per_s','ess_min','ess_min_per_s','divergent','rounds','lanes')}, b.get('ols_check',{}).get('max_abs_z_of_posterior_mean_vs_ols'))
    except Exception as e: print(f,'ERR',e); print(open(f'{O}/{f}.err').read()[-800:])
PY
