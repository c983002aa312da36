#!/bin/bash
# 2-GPU box: batched chains / device-side NUTS on ROW SHARDS (one NCCL all-reduce of the partial sums per round), the
# step-size jitter tests, and the default bench line at 2 ranks (its `ess` record now has the device-driver arm)
O=gpurun_out/r2shard; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 120 python -m pytest tests/test_device_nuts_gpu.py -x -q > $O/tests_nuts.log 2>&1; echo "nuts tests rc=$?" >> $O/tests_nuts.log; tail -3 $O/tests_nuts.log
timeout 200 python -m pytest tests/test_multigpu_gpu.py -x -q -k "2-nccl" > $O/tests_mgpu.log 2>&1; echo "mgpu tests rc=$?" >> $O/tests_mgpu.log; tail -12 $O/tests_mgpu.log
( time timeout 240 $TR --master-port 29921 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline ) > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
O='gpurun_out/r2shard'
try:
    d=json.loads(open(O+'/bench_n2.json').read().strip().splitlines()[-1]); e=d.get('ess') or {}
    print('n2', round(d['value'],1), round(d['e2e']['value'],1), d['roofline']['frac'], (d.get('parity') or {}).get('ok'))
    for k in ('b200','b200_device_driver'):
        x=e.get(k) or {}; print(k, {q: x.get(q) for q in ('wall_s','grad_evals_per_s','ess_min','ess_min_per_s','rounds','unavailable','first_draws_max_abs_diff_vs_service')})
except Exception as ex: print('ERR', ex); print(open(O+'/bench_n2.err').read()[-1500:])
PY
grep real $O/bench_n2.err
