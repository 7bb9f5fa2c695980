#!/bin/bash
# round 2, GPU call 5 (8 GPUs): the north-star target run C3 (1e8 resamples split over 8 B200, strong scaling, with the
# N-GPU == 1-GPU and ensemble-vs-RANLUX checks), the multi-GPU tests, and the weak-scaling headline line at N = 8
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=gpurun_out/r2_5
nvidia-smi --query-gpu=index,name --format=csv,noheader > ${T}_gpus.txt; wc -l ${T}_gpus.txt
timeout 300 python -m pytest -m gpu -v -s --timeout 200 tests/test_nccl_combine.py "tests/test_gpu_parity.py::test_in_process_multi_gpu_split_equals_one_gpu" > ${T}_multi_gpu_tests.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|PASSED|FAILED|SKIPPED|ERROR" ${T}_multi_gpu_tests.log | tail -8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --workload c3 --steps 3 --warmup 3 > ${T}_bench_c3_8gpu.json 2> ${T}_bench_c3_8gpu.err; echo "c3 x8 exit $?"; cut -c1-700 ${T}_bench_c3_8gpu.json; tail -3 ${T}_bench_c3_8gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 > ${T}_bench_c2_8gpu.json 2> ${T}_bench_c2_8gpu.err; echo "c2 x8 exit $?"; cut -c1-400 ${T}_bench_c2_8gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload c3 --steps 2 --warmup 3 > ${T}_bench_c3_2gpu.json 2> ${T}_bench_c3_2gpu.err; echo "c3 x2 exit $?"; cut -c1-400 ${T}_bench_c3_2gpu.json
ls -la gpurun_out | tail -8
