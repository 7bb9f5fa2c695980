#!/bin/bash
# round 2, GPU call 15: refill schedule with the two halves of a chunk on two streams (tail overlap) and the
# rare solver paths rolled; full GPU suite; C2 bench with and without the split
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=gpurun_out/r2_15
timeout 1200 python -m pytest -m gpu -q -x --timeout 400 tests > ${T}_gpu_tests.log 2>&1; echo "pytest exit $?"; tail -5 ${T}_gpu_tests.log
timeout 400 python bench.py --no-cpu-baseline > ${T}_bench_c2.json 2> ${T}_bench_c2.err; echo "bench c2 exit $?"; cut -c1-330 ${T}_bench_c2.json
FMC_NO_SPLIT=1 timeout 400 python bench.py --no-cpu-baseline > ${T}_bench_c2_nosplit.json 2> ${T}_bench_c2_nosplit.err; echo "bench c2 (no split) exit $?"; cut -c1-330 ${T}_bench_c2_nosplit.json
timeout 400 python bench.py --workload c5 --schedule refill --resamples 2000000 --steps 2 --warmup 3 --no-cpu-baseline > ${T}_bench_c5_refill.json 2> ${T}_bench_c5_refill.err; echo "bench c5 refill exit $?"; cut -c1-330 ${T}_bench_c5_refill.json
FMC_NO_SPLIT=1 timeout 400 python bench.py --workload c5 --schedule refill --resamples 2000000 --steps 2 --warmup 3 --no-cpu-baseline > ${T}_bench_c5_refill_nosplit.json 2> ${T}_bench_c5_refill_nosplit.err; echo "bench c5 refill (no split) exit $?"; cut -c1-330 ${T}_bench_c5_refill_nosplit.json
timeout 400 python bench.py --workload c5 --schedule rounds --resamples 2000000 --steps 2 --warmup 3 --no-cpu-baseline > ${T}_bench_c5_rounds.json 2> ${T}_bench_c5_rounds.err; echo "bench c5 rounds exit $?"; cut -c1-330 ${T}_bench_c5_rounds.json
