"""Mnemonic counts of selected kernels from `cuobjdump -sass stan_b200/lib/libb200glm.so` -> profiles/r2_sass_mnemonics.txt (runs without a GPU)."""
import re, subprocess, collections, sys
txt = subprocess.run(['cuobjdump', '-sass', 'stan_b200/lib/libb200glm.so'], capture_output=True, text=True).stdout
blocks = re.split(r'\n\s*Function : ', txt)[1:]
def demangle(n):
    return subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip().replace("(int)", "").replace("(bool)", "")
def _old(n):
    return subprocess.run(['cu++filt', n], capture_output=True, text=True).stdout.strip()
want = [
 ("glm_multi_kernel<0, 13, 4>", "glm_multi_kernel<bernoulli_logit, CPL=13, 4 chains> (cfg2, four chains per pass)"),
 ("glm_multi_kernel<3, 13, 4>", "glm_multi_kernel<binomial_logit, CPL=13, 4 chains>"),
 ("glm_class_kernel<0, 13, 4>", "glm_class_kernel<ordered_logistic, 13 column slots, 4> (K = 100, 5 classes)"),
 ("glm_class_kernel<1,", "glm_class_kernel<categorical_logit, ...> (first instantiation)"),
 ("glm_batched_kernel<2, 13, 0>", "glm_batched_kernel<normal_id, MBH=13, normal> (cfg3, round-2 form)"),
 ("glm_batched_kernel<3, 13, 0>", "glm_batched_kernel<binomial_logit, MBH=13, normal>"),
 ("glm_fused_kernel<0, 13>", "glm_fused_kernel<bernoulli_logit, CPL=13> (cfg2, round-2 form)"),
 ("glm_wide_kernel<0, 8, 1, 8>", "glm_wide_kernel<bernoulli_logit, 8, 1, 8> (cfg5, round-2 form)"),
 ("nuts_step_kernel", "nuts_step_kernel (device-side NUTS: tree / adaptation step, one warp per chain)"),
 ("nuts_begin_kernel", "nuts_begin_kernel"),
 ("batched_reduce_kernel", "batched_reduce_kernel"),
 ("batched_finish_kernel", "batched_finish_kernel"),
]
names = []
for b in blocks:
    mangled = b.split('\n', 1)[0].strip()
    names.append((demangle(mangled), b))
out = ["# SASS mnemonic counts (cuobjdump -sass stan_b200/lib/libb200glm.so, sm_100a) of the round-2 kernels, final build",
       "# UBLKCP = cp.async.bulk (TMA engine, 1-D bulk copy), SYNCS = mbarrier ops, DMMA = fp64 tensor MMA (mma.sync.m8n8k4.f64),",
       "# DFMA/DADD/DMUL = fp64 pipe, LDS = shared loads, MUFU = special-function unit (rcp / ex2 seeds); PREEXIT / ACQBULK = programmatic dependent launch",
       ""]
keys = ["UBLKCP","SYNCS","DMMA","DFMA","DADD","DMUL","LDS","STS","LDG","STG","SHFL","MUFU","BAR","ACQBULK","PREEXIT","ATOMG","MEMBAR","BRA","CALL"]
for pat, title in want:
    hit = [(n, b) for n, b in names if pat in n]
    if not hit:
        out.append(f"## {title}\n   (no instantiation matched {pat!r})\n"); continue
    n, b = hit[0]
    ops = collections.Counter()
    for line in b.split('\n'):
        m = re.search(r'/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
        if m: ops[m.group(1).split('.')[0]] += 1
    tot = sum(ops.values())
    out.append(f"## {title}")
    out.append(f"   {n.split('(')[0]}")
    out.append(f"   total instructions {tot}")
    out.append("   " + "  ".join(f"{k}={ops[k]}" for k in keys if ops[k]))
    out.append("   top: " + "  ".join(f"{k}={v}" for k, v in ops.most_common(12)))
    out.append("")
open('profiles/r2_sass_mnemonics.txt','w').write("\n".join(out))
print("\n".join(out))
