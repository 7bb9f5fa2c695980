#!/bin/bash
# round 2, GPU call 28: the stress model's refill kernel built with the compact (all rolled) solver
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=gpurun_out/r2_28
timeout 600 python -m pytest -m gpu -q --timeout 300 tests/test_round_schedule.py tests/test_c5_stress.py tests/test_gpu_parity.py tests/test_analytic_jacobian.py > ${T}_tests.log 2>&1; echo "pytest exit $?"; tail -3 ${T}_tests.log; grep -E "\[rounds c5" ${T}_tests.log | cut -c1-200
timeout 400 python bench.py --workload c5 --resamples 4000000 --steps 2 --warmup 3 --no-cpu-baseline > ${T}_bench_c5.json 2> ${T}_bench_c5.err; echo "bench c5 exit $?"; cut -c1-230 ${T}_bench_c5.json
timeout 200 python tools/gpu_schedule_bench.py c5mild > ${T}_schedule_c5mild.json 2> ${T}_schedule_c5mild.err; echo "c5mild exit $?"; cut -c1-160 ${T}_schedule_c5mild.err
