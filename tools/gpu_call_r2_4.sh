#!/bin/bash
# round 2, GPU call 4: conditional-graph probe, round-schedule tests, schedule comparison
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=gpurun_out/r2_4
timeout 60 tools/cond_graph_probe > ${T}_cond_graph_probe.txt 2>&1; echo "probe exit $?"; cat ${T}_cond_graph_probe.txt
timeout 600 python -m pytest -m gpu -v -s --timeout 300 tests/test_round_schedule.py > ${T}_round_tests.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|PASSED|FAILED|ERROR|Timeout|\[rounds" ${T}_round_tests.log | tail -30
timeout 400 python tools/gpu_schedule_bench.py c2 c5 c4 c1 > ${T}_schedule.json 2> ${T}_schedule.err; echo "schedule exit $?"; cut -c1-600 ${T}_schedule.err | tail -8
ls -la gpurun_out | tail -6
