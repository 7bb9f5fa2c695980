#!/bin/bash
# round 2, GPU call 23: round schedule with the residual sum of a trial point first (spec = 0; default for p > 6)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=gpurun_out/r2_23
timeout 900 python -m pytest -m gpu -v -s --timeout 300 tests/test_round_schedule.py > ${T}_tests.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|FAILED|ERROR|\[rounds" ${T}_tests.log | cut -c1-230 | tail -20
for spec in 0 1; do
FMC_ROUND_SPECULATE=$spec timeout 400 python tools/gpu_schedule_bench.py c4 c5 c2 > ${T}_schedule_spec$spec.json 2> ${T}_schedule_spec$spec.err; echo "schedule spec=$spec exit $?"; python -c "
import json; d=json.load(open('${T}_schedule_spec$spec.json'))
for k in ('c4','c5','c2'): print(k, 'rounds %.4g (frac %.3f, passes %.2f, launches %d) refill %.4g' % (d[k]['rounds']['fits_per_s'], d[k]['rounds']['roofline_frac'], d[k]['rounds']['passes_per_fit'], d[k]['rounds']['launches'], d[k]['refill']['fits_per_s']))"
done
timeout 900 python bench.py --workload c4 --steps 2 --warmup 3 > ${T}_bench_c4.json 2> ${T}_bench_c4.err; echo "bench c4 exit $?"; cut -c1-250 ${T}_bench_c4.json
