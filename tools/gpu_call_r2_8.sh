#!/bin/bash
# round 2, GPU call 8: where C5's time goes on the warp-cooperative build -- per-kernel launch list of the
# round schedule, full ncu captures of the refill kernel and of a mid-run solve / pass kernel pair
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=gpurun_out/r2_8
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"iterate_kernel" -c 1 -o ${T}_ncu_c5_refill -f python tools/gpu_profile_c2.py c5 150000 1 fd thread > ${T}_ncu_c5_refill.log 2>&1; echo "ncu c5 refill exit $?"; tail -1 ${T}_ncu_c5_refill.log
export FMC_SCHEDULE=rounds FMC_ROUNDS_NO_GRAPH=1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${T}_launches_c5_rounds.csv python tools/gpu_profile_c2.py c5 200000 1 fd thread > ${T}_launches_c5_rounds.log 2>&1; echo "launch list c5 exit $?"; tail -1 ${T}_launches_c5_rounds.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"round_(solve|pass)_kernel" --launch-skip 60 --launch-count 2 -o ${T}_ncu_c5_rounds -f python tools/gpu_profile_c2.py c5 200000 1 fd thread > ${T}_ncu_c5_rounds.log 2>&1; echo "ncu c5 rounds exit $?"; tail -1 ${T}_ncu_c5_rounds.log
ls -la gpurun_out | tail -8
