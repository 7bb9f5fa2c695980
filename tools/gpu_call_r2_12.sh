#!/bin/bash
# round 2, GPU call 12: variants of the refill kernel on C2 / C1 / C5 -- solver compiled rolled + out of line
# (as in the solve kernel), Philox-7, both; schedule identity with the rolled refill kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=gpurun_out/r2_12
V=fitter_mc_b200/variants
FMC_VARIANT_CONFIGS=c2,c1,c5 timeout 600 python tools/gpu_variant_lib_bench.py fitter_mc_b200/libfmc_b200.so $V/lib_rolled.so $V/lib_philox7.so $V/lib_rolled_philox7.so fitter_mc_b200/libfmc_b200.so > ${T}_variants.json 2> ${T}_variants.err; echo "variants exit $?"; python -c "
import json; d=json.load(open('${T}_variants.json'))
for k,v in d.items(): print(k, {c: '%.4g' % r['fits_per_s'] for c,r in v.items()} if isinstance(v,dict) else v[-300:])"
FMC_B200_LIB=$PWD/$V/lib_rolled.so timeout 600 python -m pytest -m gpu -v -s --timeout 300 tests/test_round_schedule.py tests/test_gpu_parity.py > ${T}_tests_rolled.log 2>&1; echo "pytest(rolled) exit $?"; grep -E "passed|failed|FAILED|ERROR|\[rounds|\[parity" ${T}_tests_rolled.log | cut -c1-220 | tail -24
