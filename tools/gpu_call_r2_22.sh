#!/bin/bash
# round 2, GPU call 22: lmstep as a warp-cooperative state machine (the More-Hebden iterations of the lanes of a warp run together)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=gpurun_out/r2_22
V=fitter_mc_b200/variants
FMC_VARIANT_CONFIGS=c2,c1,c5 timeout 900 python tools/gpu_variant_lib_bench.py $V/lib_base.so fitter_mc_b200/libfmc_b200.so > ${T}_variants.json 2> ${T}_variants.err; echo "variants exit $?"; python -c "
import json; d=json.load(open('${T}_variants.json'))
for k,v in d.items(): print(k, {c: '%.4g' % r['fits_per_s'] for c,r in v.items()} if isinstance(v,dict) else v[-300:])"
timeout 300 python tools/gpu_schedule_bench.py c5 c4 > ${T}_schedule.json 2> ${T}_schedule.err; echo "schedule exit $?"; python -c "
import json; d=json.load(open('${T}_schedule.json'))
for k in ('c5','c4'): print(k, 'rounds %.4g refill %.4g' % (d[k]['rounds']['fits_per_s'], d[k]['refill']['fits_per_s']))"
timeout 900 python -m pytest -m gpu -q -x --timeout 300 tests > ${T}_gpu_tests.log 2>&1; echo "pytest exit $?"; tail -3 ${T}_gpu_tests.log
