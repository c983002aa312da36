#!/bin/bash
# N-GPU validation pass: the row-sharded parity tests, then bench.py at N ranks with and without programmatic
# dependent launch (B200GLM_NO_PDL=1).  Usage: gpurun --gpus N -- bash tools/mgpu_check.sh N
N=${1:-2}
OUT=gpurun_out/mgpu$N
mkdir -p $OUT
(time timeout 300 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q) > $OUT/tests.log 2>&1
tail -3 $OUT/tests.log
port=29900
for pdl in 0 1 0 1; do
  port=$((port + 1))
  B200GLM_NO_PDL=$pdl timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $port bench.py --gpus $N --steps 400 --warmup 20 --no-cpu-baseline > $OUT/cfg2_nopdl$pdl.$port.json 2> $OUT/err.$port
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/cfg2_nopdl$pdl.$port.json").read().strip().splitlines()[-1])
    print("NO_PDL=$pdl n=", d["n_gpus"], "value=%.1f ms=%.4f e2e=%.1f frac=%.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"]))
except Exception as e:
    print("NO_PDL=$pdl failed", e); print(open("$OUT/err.$port").read()[-1500:])
PY
done
