#!/bin/bash
# round 2, GPU call 7: warp-cooperative solver state machine -- round-schedule tests, parity, schedule comparison
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=gpurun_out/r2_7
timeout 600 python -m pytest -m gpu -v -s --timeout 300 tests/test_round_schedule.py tests/test_gpu_parity.py tests/test_c5_stress.py > ${T}_tests.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|FAILED|ERROR|Timeout|\[rounds|\[parity" ${T}_tests.log | tail -30
timeout 400 python tools/gpu_schedule_bench.py c2 c5 c4 c1 > ${T}_schedule.json 2> ${T}_schedule.err; echo "schedule exit $?"; cut -c1-250 ${T}_schedule.err | tail -8
ls -la gpurun_out | tail -4
