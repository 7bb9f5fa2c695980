#!/bin/bash
# round 2, GPU call 6: headline bench with the single-call-site refill kernel; per-kernel breakdown of the round schedule on C5
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=gpurun_out/r2_6
timeout 300 python bench.py > ${T}_bench_c2.json 2> ${T}_bench_c2.err; echo "bench c2 exit $?"; cut -c1-300 ${T}_bench_c2.json
export FMC_SCHEDULE=rounds FMC_ROUNDS_NO_GRAPH=1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${T}_launches_c5_rounds.csv python tools/gpu_profile_c2.py c5 200000 1 fd thread > ${T}_launches_c5_rounds.log 2>&1; echo "launch list c5 exit $?"; tail -1 ${T}_launches_c5_rounds.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"round_(solve|pass)_kernel" --launch-skip 200 --launch-count 4 -o ${T}_ncu_c5_rounds -f python tools/gpu_profile_c2.py c5 200000 1 fd thread > ${T}_ncu_c5_rounds.log 2>&1; echo "ncu c5 rounds exit $?"; tail -1 ${T}_ncu_c5_rounds.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${T}_launches_c2_rounds.csv python tools/gpu_profile_c2.py c2 1000000 1 fd thread > ${T}_launches_c2_rounds.log 2>&1; echo "launch list c2 exit $?"; tail -1 ${T}_launches_c2_rounds.log
ls -la gpurun_out | tail -8
