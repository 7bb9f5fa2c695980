#!/bin/bash
# round 2, GPU call 29: final tree -- full GPU suite, smoke, default bench line, C5 bench line at 1e7 resamples per step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=gpurun_out/r2_29
timeout 1200 python -m pytest -m gpu -q --timeout 400 tests > ${T}_gpu_tests.log 2>&1; echo "pytest exit $?"; tail -3 ${T}_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${T}_smoke.log 2>&1; echo "smoke exit $?"; tail -1 ${T}_smoke.log
timeout 500 python bench.py > ${T}_bench_c2.json 2> ${T}_bench_c2.err; echo "bench c2 exit $?"; cut -c1-200 ${T}_bench_c2.json
timeout 600 python bench.py --workload c5 --steps 2 --warmup 3 > ${T}_bench_c5.json 2> ${T}_bench_c5.err; echo "bench c5 exit $?"; cut -c1-230 ${T}_bench_c5.json
