#!/bin/bash
# last verification pass of the round on a 2-GPU box: the whole GPU suite (incl. the 2-rank tests), the default bench
# line as the driver runs it, and the smoke test.  Every command has its own timeout.
O=gpurun_out/r2final; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log; tail -6 $O/tests.log
( time timeout 600 python bench.py --steps 20 --warmup 5 ) > $O/bench_default.json 2> $O/bench_default.err; echo "default rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29877 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; echo "n2 rc=$?"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python - <<'PY'
import json
O='gpurun_out/r2final'
for f in ['bench_default','bench_n2']:
    try:
        d=json.loads(open(f'{O}/{f}.json').read().strip().splitlines()[-1])
        e=d.get('ess') or {}
        print(f, round(d['value'],2), round(d['ms_per_step'],4), round(d['e2e']['value'],1), d['roofline']['frac'], (d.get('parity') or {}).get('max_rel_err'), (e.get('b200') or {}).get('ess_min_per_s'), json.dumps(e.get('config1_both_arms'))[:600])
    except Exception as ex: print(f,'ERR',ex)
PY
grep real $O/bench_default.err
