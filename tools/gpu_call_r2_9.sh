#!/bin/bash
# round 2, GPU call 9: solve kernel with the state in shared memory, rolled solver, out-of-line helpers
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=gpurun_out/r2_9
timeout 600 python -m pytest -m gpu -v -s --timeout 300 tests/test_round_schedule.py tests/test_c5_stress.py tests/test_warp_mapping.py > ${T}_tests.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|FAILED|ERROR|Timeout|\[rounds|\[parity" ${T}_tests.log | tail -30
timeout 400 python tools/gpu_schedule_bench.py c5 c2 c4 c1 > ${T}_schedule.json 2> ${T}_schedule.err; echo "schedule exit $?"; cut -c1-250 ${T}_schedule.err | tail -8
export FMC_SCHEDULE=rounds FMC_ROUNDS_NO_GRAPH=1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${T}_launches_c5_rounds.csv python tools/gpu_profile_c2.py c5 200000 1 fd thread > ${T}_launches_c5_rounds.log 2>&1; echo "launch list c5 exit $?"; tail -1 ${T}_launches_c5_rounds.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"round_(solve|pass)_kernel" --launch-skip 60 --launch-count 2 -o ${T}_ncu_c5_rounds -f python tools/gpu_profile_c2.py c5 200000 1 fd thread > ${T}_ncu_c5_rounds.log 2>&1; echo "ncu c5 rounds exit $?"; tail -1 ${T}_ncu_c5_rounds.log
ls -la gpurun_out | tail -8
