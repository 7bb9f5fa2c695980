#!/bin/bash
# round 2, GPU call 25: the full GPU suite, smoke() and the default bench line on the tree with the residual-first rounds
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=gpurun_out/r2_25
timeout 1200 python -m pytest -m gpu -q --timeout 400 tests > ${T}_gpu_tests.log 2>&1; echo "pytest exit $?"; tail -4 ${T}_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${T}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 ${T}_smoke.log
timeout 500 python bench.py > ${T}_bench_c2.json 2> ${T}_bench_c2.err; echo "bench c2 exit $?"; cut -c1-200 ${T}_bench_c2.json
