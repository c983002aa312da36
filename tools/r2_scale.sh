#!/bin/bash
# Round-2 multi-GPU pass (run under `gpurun --gpus N`, N = 2 for the dry run, 8 for the record): the 2/8-rank parity
# tests, the per-phase timeline of the sharded step, then bench.py at 1, 2, 4, 8 GPUs for config 2 (strong scaling,
# the metric's config; each line carries parity + ESS/s), --balance, config 4 (poisson + groups, fused group path)
# and config 5 at HBM scale (streamed build, as many rows per GPU as fit).  Every command has its own timeout.
OUT=gpurun_out/r2scale
mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l)
echo "gpus: $NG" > $OUT/host.log
(time timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q) > $OUT/tests_mgpu.log 2>&1
tail -3 $OUT/tests_mgpu.log
port=29610
tr() {  # torchrun n script args...
  local n=$1; shift
  port=$((port + 1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port "$@"
}
run() {  # name, n, timeout, args...
  local name=$1 n=$2 to=$3; shift 3
  if [ "$n" -gt "$NG" ]; then return; fi
  if [ "$n" -eq 1 ]; then
    timeout $to python bench.py --gpus 1 "$@" > $OUT/${name}_n1.json 2> $OUT/${name}_n1.err
  else
    timeout $to python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
      --master-port $((port = port + 1)) bench.py --gpus $n "$@" > $OUT/${name}_n$n.json 2> $OUT/${name}_n$n.err
  fi
  echo "$name n=$n rc=$?" >> $OUT/host.log
}
# per-phase timeline of the sharded step at the N-GPU shard size of config 2
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29601 \
  tools/timeline_probe.py --rows 10000000 --out $OUT/timeline_n${NG}.json > $OUT/timeline_n${NG}.log 2>&1
for n in 1 2 4 8; do run cfg2 $n 400 --steps 400 --warmup 20 --cpu-evals 3; done
run cfg2balance $NG 300 --steps 400 --warmup 20 --balance --no-cpu-baseline --ess-iters 0
run cfg2nccl $NG 300 --steps 400 --warmup 20 --collective nccl --no-cpu-baseline --ess-iters 0
for n in 1 $NG; do run cfg4 $n 400 --config 4 --steps 100 --warmup 10 --no-cpu-baseline; done
run cfg5 $NG 600 --config 5 --streamed --rows 0 --steps 10 --warmup 3 --no-cpu-baseline
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2scale/*_n*.json')):
    if 'timeline' in f: continue
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = (d.get('ess') or {}).get('b200') or {}
        print(f.split('/')[-1], 'n=', d['n_gpus'], 'value=%.1f ms=%.4f e2e=%.1f frac=%.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']),
              'parity=', (d.get('parity') or {}).get('max_rel_err'), 'ess/s=', e.get('ess_min_per_s'), 'rows/gpu=', d['config'].get('rows_per_gpu'))
    except Exception as ex:
        print(f, 'ERR', ex)
        try: print(open(f.replace('.json', '.err')).read()[-800:])
        except Exception: pass
try:
    for ln in open(glob.glob('gpurun_out/r2scale/timeline_n*.json')[0]):
        d = json.loads(ln)
        print('rank', d['rank'], round(d['us_per_step_events_plain'], 2), {k: round(v, 2) for k, v in d['phases_us'].items()})
except Exception as ex:
    print('timeline ERR', ex)
PY
