#!/bin/bash
# round 2, GPU call 14: which further solver routines to compile rolled in the refill kernel
# (A: adopt_jacobian, factor, build_h, hess_times_step; B: the head of lmstep; AB: both)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=gpurun_out/r2_14
V=fitter_mc_b200/variants
FMC_VARIANT_CONFIGS=c2,c1,c5 timeout 900 python tools/gpu_variant_lib_bench.py $V/lib_rare.so $V/lib_rare_A.so $V/lib_rare_B.so $V/lib_rare_AB.so > ${T}_variants.json 2> ${T}_variants.err; echo "variants exit $?"; python -c "
import json; d=json.load(open('${T}_variants.json'))
for k,v in d.items(): print(k, {c: '%.4g' % r['fits_per_s'] for c,r in v.items()} if isinstance(v,dict) else v[-300:])"
