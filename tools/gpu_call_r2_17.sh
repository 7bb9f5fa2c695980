#!/bin/bash
# round 2, GPU call 17: v2norm's overflow fallback out of line everywhere; ncu of the C4 round kernels (host-driven rounds:
# ncu cannot profile kernels inside a conditional graph node)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=gpurun_out/r2_17
FMC_VARIANT_CONFIGS=c2,c1,c5 timeout 600 python tools/gpu_variant_lib_bench.py fitter_mc_b200/libfmc_b200.so > ${T}_variants.json 2> ${T}_variants.err; echo "variants exit $?"; python -c "
import json; d=json.load(open('${T}_variants.json'))
for k,v in d.items(): print(k, {c: '%.4g' % r['fits_per_s'] for c,r in v.items()} if isinstance(v,dict) else v[-300:])"
timeout 900 python -m pytest -m gpu -q -x --timeout 400 tests > ${T}_gpu_tests.log 2>&1; echo "pytest exit $?"; tail -3 ${T}_gpu_tests.log
FMC_ROUNDS_NO_GRAPH=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"round_(solve|pass)_kernel" --launch-skip 8 --launch-count 2 -o ${T}_ncu_c4 -f python tools/gpu_profile_c2.py c4 100000 1 fd thread > ${T}_ncu_c4.log 2>&1; echo "ncu c4 exit $?"; tail -1 ${T}_ncu_c4.log
timeout 300 python bench.py --no-cpu-baseline > ${T}_bench_c2.json 2> ${T}_bench_c2.err; echo "bench c2 exit $?"; cut -c1-200 ${T}_bench_c2.json
