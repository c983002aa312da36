#!/bin/bash
O=gpurun_out/r2c13; mkdir -p $O
export NCU_SUMMARY_DIR=$PWD/$O
sum() { # target name N K family G kernel
  timeout 300 ncu --set full --clock-control none -k regex:glm_ --launch-skip 4 --launch-count 2 -f -o /tmp/ncu_$1 python tools/ncu_target.py $1 > $O/ncu_$1.log 2>&1
  timeout 120 python profiles/summarize_ncu.py /tmp/ncu_$1.ncu-rep $2 $3 $4 $5 $6 $7 > /dev/null 2>> $O/ncu_$1.log
  rm -f /tmp/ncu_$1.ncu-rep; echo "ncu $1 done"
}
sum cfg2 r2_glm_fused_bernoulli_N10M_K100 10000000 100 bernoulli_logit 0 glm_fused_kernel
sum cfg2shard r2_glm_fused_bernoulli_N1250k_K100_shard 1250000 100 bernoulli_logit 0 glm_fused_kernel
sum cfg4shard r2_glm_fused_poisson_groups_N6250k_K50_shard 6250000 50 poisson_log 1000 glm_fused_kernel
sum wide r2_glm_wide_bernoulli_N1M_K1000 1000000 1000 bernoulli_logit 0 glm_wide_kernel
sum cfg3 r2_glm_batched_normal_N1M_K200_C1024 1000000 200 normal_id 0 glm_batched_kernel
sum ordlog r2_glm_class_ordered_logistic_N10M_K100_C5 10000000 100 ordered_logistic 0 glm_class_kernel
sum catlog r2_glm_class_categorical_logit_N10M_K100_C4 10000000 100 categorical_logit 0 glm_class_kernel
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r2_launches_bench_N10M_K100.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity --ess-iters 0 > $O/launches_bench.log 2>&1
timeout 300 python -m pytest tests/test_stan_dropin_gpu.py -m gpu -x -q -k "batched" 2>&1 | tail -2
( time timeout 480 python bench_nuts.py --config 3 --chains 1024 --warmup 150 --samples 100 ) > $O/nuts_cfg3_150_100.json 2> $O/nuts_cfg3_150_100.err; echo "nuts cfg3 rc=$?"
cut -c1-2500 $O/nuts_cfg3_150_100.json; tail -4 $O/nuts_cfg3_150_100.err
du -sh $O; ls $O
