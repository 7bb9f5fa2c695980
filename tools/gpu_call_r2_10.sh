#!/bin/bash
# round 2, GPU call 10: solve kernel with 8 slots per warp (and the 4 / 16 variants), relaxed identity tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=gpurun_out/r2_10
timeout 600 python -m pytest -m gpu -v -s --timeout 300 tests/test_round_schedule.py > ${T}_tests.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|FAILED|ERROR|Timeout|\[rounds|\[parity" ${T}_tests.log | tail -30
timeout 400 python tools/gpu_schedule_bench.py c5 c2 c4 c1 > ${T}_schedule.json 2> ${T}_schedule.err; echo "schedule exit $?"; cut -c1-120 ${T}_schedule.err | tail -8
for L in 4 16; do
FMC_B200_LIB=$PWD/fitter_mc_b200/variants/libfmc_b200_L$L.so timeout 300 python tools/gpu_schedule_bench.py c5 c2 > ${T}_schedule_L$L.json 2> ${T}_schedule_L$L.err; echo "schedule L=$L exit $?"; cut -c1-120 ${T}_schedule_L$L.err | tail -4
done
export FMC_SCHEDULE=rounds FMC_ROUNDS_NO_GRAPH=1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${T}_launches_c5_rounds.csv python tools/gpu_profile_c2.py c5 200000 1 fd thread > ${T}_launches_c5_rounds.log 2>&1; echo "launch list c5 exit $?"; tail -1 ${T}_launches_c5_rounds.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"round_(solve|pass)_kernel" --launch-skip 60 --launch-count 2 -o ${T}_ncu_c5_rounds -f python tools/gpu_profile_c2.py c5 200000 1 fd thread > ${T}_ncu_c5_rounds.log 2>&1; echo "ncu c5 rounds exit $?"; tail -1 ${T}_ncu_c5_rounds.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${T}_launches_c2_rounds.csv python tools/gpu_profile_c2.py c2 1000000 1 fd thread > ${T}_launches_c2_rounds.log 2>&1; echo "launch list c2 exit $?"; tail -1 ${T}_launches_c2_rounds.log
ls -la gpurun_out | tail -8
