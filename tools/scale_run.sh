#!/bin/bash
# Multi-GPU measurement pass (run under `gpurun --gpus 8`): the 2/8-rank parity tests, then bench.py at
# N = 1, 2, 4, 8 for configs 2 (strong scaling, the metric's config), 4 (poisson + groups) and 5 (weak).
# Every JSON line lands in gpurun_out/scale/; copy what is to be judged into profiles/.
OUT=gpurun_out/scale
mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l)
echo "gpus: $NG" > $OUT/host.log
nvidia-smi topo -m >> $OUT/host.log 2>&1
(time timeout 300 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q) > $OUT/tests_mgpu.log 2>&1
port=29810
run() {  # name, n, args...
  local name=$1 n=$2; shift 2
  port=$((port + 1))
  if [ "$n" -gt "$NG" ]; then return; fi
  if [ "$n" -eq 1 ]; then
    timeout 150 python bench.py --gpus 1 "$@" > $OUT/${name}_n1.json 2> $OUT/${name}_n1.err
  else
    timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
      --master-port $port bench.py --gpus $n "$@" > $OUT/${name}_n$n.json 2> $OUT/${name}_n$n.err
  fi
  echo "$name n=$n rc=$?" >> $OUT/host.log
}
for n in 1 2 4 8; do run cfg2 $n --steps 200 --warmup 10 --no-cpu-baseline; done
for n in 8; do run cfg2nccl $n --steps 200 --warmup 10 --collective nccl; done
for n in 4 8; do run cfg4 $n --config 4 --steps 100 --warmup 10; done
for n in 8; do run cfg5 $n --config 5 --steps 30 --warmup 5; done
grep -h '"metric"' $OUT/*.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['config']['workload'][:40], 'n=', d['n_gpus'], 'value=%.1f' % d['value'], 'e2e=%.1f' % d['e2e']['value'], 'frac=%.3f' % d['roofline']['frac'])
"
tail -3 $OUT/tests_mgpu.log
