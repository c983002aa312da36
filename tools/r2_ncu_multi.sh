#!/bin/bash
O=gpurun_out/r2ncu_multi; mkdir -p $O
timeout 400 ncu --set full --clock-control none --import-source on -k regex:glm_multi --launch-skip 2 --launch-count 1 -f -o $O/multi4 python tools/ncu_target.py multi4 > $O/ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $O/ncu.log
NCU_SUMMARY_DIR=$O python profiles/summarize_ncu.py $O/multi4.ncu-rep r2_glm_multi_bernoulli_N4M_K100_C4 4000000 100 bernoulli_logit 0 glm_multi > $O/sum.log 2>&1; tail -3 $O/sum.log
ls -la $O
