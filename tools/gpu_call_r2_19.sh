#!/bin/bash
# round 2, GPU call 19: block size of the refill kernel re-measured on the final solver flavour (256 shipped, 320, 384)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=gpurun_out/r2_19
V=fitter_mc_b200/variants
FMC_VARIANT_CONFIGS=c2,c1,c5 timeout 900 python tools/gpu_variant_lib_bench.py fitter_mc_b200/libfmc_b200.so $V/lib_t320.so $V/lib_t384.so > ${T}_variants.json 2> ${T}_variants.err; echo "variants exit $?"; python -c "
import json; d=json.load(open('${T}_variants.json'))
for k,v in d.items(): print(k, {c: '%.4g' % r['fits_per_s'] for c,r in v.items()} if isinstance(v,dict) else v[-300:])"
