#!/bin/bash
# round 2, GPU call 13: refill kernel with only the rare solver paths (gqtstp, lsvmin, More-Hebden loop)
# rolled and out of line; the same + Philox-7; solve kernel with unrolled loops; fixed round tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=gpurun_out/r2_13
V=fitter_mc_b200/variants
FMC_VARIANT_CONFIGS=c2,c1,c5 timeout 600 python tools/gpu_variant_lib_bench.py fitter_mc_b200/libfmc_b200.so $V/lib_rare.so $V/lib_rare_philox7.so > ${T}_variants.json 2> ${T}_variants.err; echo "variants exit $?"; python -c "
import json; d=json.load(open('${T}_variants.json'))
for k,v in d.items(): print(k, {c: '%.4g' % r['fits_per_s'] for c,r in v.items()} if isinstance(v,dict) else v[-300:])"
FMC_B200_LIB=$PWD/$V/lib_solve_unrolled.so timeout 300 python tools/gpu_schedule_bench.py c5 c2 > ${T}_schedule_solve_unrolled.json 2> ${T}_schedule_solve_unrolled.err; echo "schedule (solve unrolled) exit $?"; cut -c1-120 ${T}_schedule_solve_unrolled.err | tail -4
python -c "
import json; d=json.load(open('${T}_schedule_solve_unrolled.json'))
for k in ('c5','c2'): print(k, 'rounds %.4g refill %.4g' % (d[k]['rounds']['fits_per_s'], d[k]['refill']['fits_per_s']))"
FMC_B200_LIB=$PWD/$V/lib_rare.so timeout 600 python -m pytest -m gpu -v -s --timeout 300 tests/test_round_schedule.py tests/test_c5_stress.py > ${T}_tests_rare.log 2>&1; echo "pytest(rare) exit $?"; grep -E "passed|failed|FAILED|ERROR|\[rounds|\[c5 stress cuda" ${T}_tests_rare.log | cut -c1-250 | tail -24
timeout 600 python -m pytest -m gpu -q --timeout 300 tests/test_round_schedule.py tests/test_c5_stress.py > ${T}_tests.log 2>&1; echo "pytest(default) exit $?"; tail -4 ${T}_tests.log
