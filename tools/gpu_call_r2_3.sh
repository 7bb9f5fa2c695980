#!/bin/bash
# round 2, GPU call 3: full GPU suite, mapping decision with the rolled warp solver, ncu of the C5 and C4 thread kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=gpurun_out/r2_3
timeout 900 python -m pytest -m gpu -v --timeout 200 tests > ${T}_gpu_tests.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|FAILED|ERROR|Timeout" ${T}_gpu_tests.log | tail -20
timeout 200 python tools/gpu_mapping_bench.py c4 > ${T}_mapping.json 2> ${T}_mapping.err; echo "mapping exit $?"; tail -1 ${T}_mapping.err | cut -c1-700
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"iterate_kernel" -c 1 -o ${T}_ncu_c5 -f python tools/gpu_profile_c2.py c5 150000 1 fd thread > ${T}_ncu_c5.log 2>&1; echo "ncu c5 exit $?"; tail -2 ${T}_ncu_c5.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"iterate_kernel" -c 1 -o ${T}_ncu_c4 -f python tools/gpu_profile_c2.py c4 60000 1 fd thread > ${T}_ncu_c4.log 2>&1; echo "ncu c4 exit $?"; tail -2 ${T}_ncu_c4.log
ls -la gpurun_out | tail -8
