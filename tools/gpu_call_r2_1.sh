#!/bin/bash
# round 2, GPU call 1: GPU test-suite (warp path un-skipped), mapping decision by measurement, bench, trial1
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_1_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r2_1_gpu_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_1_gpu_tests.log
tail -5 gpurun_out/r2_1_gpu_tests.log
timeout 900 python tools/gpu_mapping_bench.py c4 c2 c5mild c5 c1 > gpurun_out/r2_1_mapping.json 2> gpurun_out/r2_1_mapping.err; echo "mapping exit $?"
tail -3 gpurun_out/r2_1_mapping.err
timeout 600 python bench.py > gpurun_out/r2_1_bench.json 2> gpurun_out/r2_1_bench.err; echo "bench exit $?"
cat gpurun_out/r2_1_bench.json | cut -c1-600
FMC_SHARED_FIRST_TRIAL=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_1_bench_trial1.json 2> gpurun_out/r2_1_bench_trial1.err; echo "bench trial1 exit $?"
cat gpurun_out/r2_1_bench_trial1.json | cut -c1-300
