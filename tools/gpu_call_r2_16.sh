#!/bin/bash
# round 2, GPU call 16: evidence on the final kernels -- bench lines, ncu launch list of the bench command,
# full ncu captures of the dominant kernels of C2 (inside bench.py), C4 (round schedule) and C5 (refill)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=gpurun_out/r2_16
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > ${T}_bench_plain.json 2> ${T}_bench_plain.err; echo "bench plain exit $?"; cut -c1-200 ${T}_bench_plain.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${T}_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > ${T}_bench_launches.log 2>&1; echo "launch list exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"iterate_kernel" --launch-skip 6 --launch-count 2 -o ${T}_ncu_c2 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > ${T}_ncu_c2.log 2>&1; echo "ncu c2 exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"round_(solve|pass)_kernel" --launch-skip 8 --launch-count 2 -o ${T}_ncu_c4 -f python tools/gpu_profile_c2.py c4 100000 1 fd thread > ${T}_ncu_c4.log 2>&1; echo "ncu c4 exit $?"; tail -1 ${T}_ncu_c4.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"iterate_kernel" -c 1 -o ${T}_ncu_c5 -f python tools/gpu_profile_c2.py c5 150000 1 fd thread > ${T}_ncu_c5.log 2>&1; echo "ncu c5 exit $?"; tail -1 ${T}_ncu_c5.log
timeout 500 python bench.py > ${T}_bench_c2.json 2> ${T}_bench_c2.err; echo "bench c2 exit $?"; cut -c1-200 ${T}_bench_c2.json
timeout 600 python bench.py --workload c5 --steps 2 --warmup 3 > ${T}_bench_c5.json 2> ${T}_bench_c5.err; echo "bench c5 exit $?"; cut -c1-200 ${T}_bench_c5.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > ${T}_bench_reference.json 2> ${T}_bench_reference.err; echo "bench reference exit $?"; cut -c1-200 ${T}_bench_reference.json
ls -la gpurun_out | tail -14
