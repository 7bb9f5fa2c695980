#!/bin/bash
# round 2, GPU call 2: mapping decision (C4, C5), targeted GPU tests, bench lines, trial1, ncu captures
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=gpurun_out/r2_2
timeout 200 python tools/gpu_mapping_bench.py c4 c5 > ${T}_mapping.json 2> ${T}_mapping.err; echo "mapping exit $?"; tail -2 ${T}_mapping.err | cut -c1-900
timeout 700 python -m pytest -m gpu -v -s --timeout 150 tests/test_warp_mapping.py tests/test_c5_stress.py tests/test_gpu_parity.py tests/test_gpu_rng.py tests/test_nccl_combine.py > ${T}_gpu_tests.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|PASSED|FAILED|ERROR|Timeout" ${T}_gpu_tests.log | tail -40
timeout 300 python bench.py > ${T}_bench_c2.json 2> ${T}_bench_c2.err; echo "bench c2 exit $?"; cut -c1-400 ${T}_bench_c2.json
FMC_SHARED_FIRST_TRIAL=1 timeout 120 python bench.py --no-cpu-baseline > ${T}_bench_c2_trial1.json 2> ${T}_bench_c2_trial1.err; echo "bench trial1 exit $?"; cut -c1-200 ${T}_bench_c2_trial1.json
timeout 200 python bench.py --workload c4 --mapping warp --steps 2 --warmup 3 --resamples 400000 --no-cpu-baseline > ${T}_bench_c4_warp.json 2> ${T}_bench_c4_warp.err; echo "bench c4 exit $?"; cut -c1-300 ${T}_bench_c4_warp.json
timeout 200 python bench.py --workload c5 --steps 2 --warmup 3 --resamples 1000000 --no-cpu-baseline > ${T}_bench_c5.json 2> ${T}_bench_c5.err; echo "bench c5 exit $?"; cut -c1-300 ${T}_bench_c5.json
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"iterate_kernel" -c 2 -o ${T}_ncu_c2 -f python tools/gpu_profile_c2.py c2 262144 1 fd thread > ${T}_ncu_c2.log 2>&1; echo "ncu c2 exit $?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"warp_fit_kernel" -c 1 -o ${T}_ncu_c4 -f python tools/gpu_profile_c2.py c4 20000 1 fd warp > ${T}_ncu_c4.log 2>&1; echo "ncu c4 exit $?"
ls -la gpurun_out | tail -15
