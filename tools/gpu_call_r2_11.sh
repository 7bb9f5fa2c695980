#!/bin/bash
# round 2, GPU call 11: full GPU suite on the tree with the shared-memory solve kernel and the automatic
# schedule; bench lines of C2 (default), C4 and C5 at their BASELINE sizes (1e7 resamples per step)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=gpurun_out/r2_11
timeout 1200 python -m pytest -m gpu -q -x --timeout 400 tests > ${T}_gpu_tests.log 2>&1; echo "pytest exit $?"; tail -5 ${T}_gpu_tests.log
timeout 400 python bench.py > ${T}_bench_c2.json 2> ${T}_bench_c2.err; echo "bench c2 exit $?"; cut -c1-400 ${T}_bench_c2.json
timeout 600 python bench.py --workload c4 --steps 2 --warmup 3 > ${T}_bench_c4.json 2> ${T}_bench_c4.err; echo "bench c4 exit $?"; cut -c1-400 ${T}_bench_c4.json
timeout 600 python bench.py --workload c5 --schedule rounds --steps 2 --warmup 3 > ${T}_bench_c5.json 2> ${T}_bench_c5.err; echo "bench c5 exit $?"; cut -c1-400 ${T}_bench_c5.json
ls -la gpurun_out | tail -8
