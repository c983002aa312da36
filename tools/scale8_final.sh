#!/bin/bash
# final 8-GPU pass: cfg2 strong scaling at N = 8 with / without programmatic dependent launch, and N = 4
OUT=gpurun_out/scale8b
mkdir -p $OUT
port=29950
run() {  # tag n nopdl
  port=$((port + 1))
  B200GLM_NO_PDL=$3 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 \
    --master-port $port bench.py --gpus $2 --steps 400 --warmup 20 --no-cpu-baseline > $OUT/$1.json 2> $OUT/$1.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/$1.json").read().strip().splitlines()[-1])
    print("$1 n=", d["n_gpus"], "value=%.1f ms=%.4f e2e=%.1f frac=%.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"]))
except Exception as e:
    print("$1 failed", e); print(open("$OUT/$1.err").read()[-1500:])
PY
}
run cfg2_n8_pdl 8 0
run cfg2_n8_nopdl 8 1
run cfg2_n8_pdl_b 8 0
run cfg2_n4_pdl 4 0
timeout 100 python bench.py --gpus 1 --steps 400 --warmup 20 --no-cpu-baseline > $OUT/cfg2_n1.json 2> $OUT/cfg2_n1.err
python -c "
import json; d=json.loads(open('$OUT/cfg2_n1.json').read().strip().splitlines()[-1]); print('n=1 value=%.1f ms=%.4f' % (d['value'], d['ms_per_step']))"
