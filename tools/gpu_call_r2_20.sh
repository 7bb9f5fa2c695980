#!/bin/bash
# round 2, GPU call 20: prefetch of the solver's local frame before the last rows of a pass (refill kernel)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=gpurun_out/r2_20
V=fitter_mc_b200/variants
FMC_VARIANT_CONFIGS=c2,c1,c5 timeout 900 python tools/gpu_variant_lib_bench.py $V/lib_base.so fitter_mc_b200/libfmc_b200.so $V/lib_base.so fitter_mc_b200/libfmc_b200.so > ${T}_variants.json 2> ${T}_variants.err; echo "variants exit $?"; python -c "
import json; d=json.load(open('${T}_variants.json'))
for k,v in d.items(): print(k, {c: '%.4g' % r['fits_per_s'] for c,r in v.items()} if isinstance(v,dict) else v[-300:])"
timeout 300 python bench.py --no-cpu-baseline > ${T}_bench_c2.json 2> ${T}_bench_c2.err; echo "bench c2 exit $?"; cut -c1-200 ${T}_bench_c2.json
timeout 600 python -m pytest -m gpu -q -x --timeout 300 tests/test_round_schedule.py tests/test_gpu_parity.py > ${T}_tests.log 2>&1; echo "pytest exit $?"; tail -3 ${T}_tests.log
