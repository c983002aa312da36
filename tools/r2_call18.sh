#!/bin/bash
O=gpurun_out/r2c18; mkdir -p $O
( time timeout 420 python bench_nuts.py --config 3 --rows 100000 --chains 1024 --warmup 400 --samples 100 ) > $O/nuts_cfg3_N100k_400_100.json 2> $O/nuts_cfg3_N100k.err; echo "rc=$?"
cut -c1-3000 $O/nuts_cfg3_N100k_400_100.json; tail -4 $O/nuts_cfg3_N100k.err
