#!/bin/bash
# round 2, call 1 (2 GPUs): GPU suite incl. the 2-rank tests, timelines at 1 and 2 GPUs, bench at N=1 and N=2
set -x
O=gpurun_out/r2c1; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
timeout 300 python tools/timeline_probe.py --rows 1250000 --out $O/tl_n1_1250k.json > $O/tl_n1_1250k.log 2>&1
timeout 300 python tools/timeline_probe.py --rows 10000000 --reps 50 --out $O/tl_n1_10M.json > $O/tl_n1_10M.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 tools/timeline_probe.py --rows 2500000 --out $O/tl_n2_2500k.json > $O/tl_n2_2500k.log 2>&1
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err
timeout 300 python bench.py --rows 1250000 --steps 500 --warmup 20 --no-cpu-baseline > $O/bench_n1_1250k.json 2> $O/bench_n1_1250k.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29812 bench.py --gpus 2 --steps 300 --warmup 20 > $O/bench_n2.json 2> $O/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29813 bench.py --gpus 2 --rows 2500000 --steps 1000 --warmup 20 > $O/bench_n2_2500k.json 2> $O/bench_n2_2500k.err
tail -3 $O/tests.log; cat $O/tl_n1_1250k.json | head -c 3000; cat $O/bench_n1.json $O/bench_n2.json | cut -c1-400
